"""API call logging / tensor dump / replay.  Parity: reference flashinfer/api_logging.py (@flashinfer_api :711-1075,
dumps :1346-1700, replay_from_dump :2364, replay_sequence :2448).

Environment:
  FLASHINFER_LOGLEVEL   0 off (zero overhead: the decorator returns the function unchanged), 1 names,
                        3 names + tensor metadata, 5 additionally min/max/mean/NaN/Inf statistics (computed by the
                        native ``tensor_stats`` kernel; skipped while a CUDA graph is being captured)
  FLASHINFER_LOGDEST    stdout | stderr | <file path, %i = pid>
  FLASHINFER_DUMP_DIR   when set (and level >= 3) inputs are saved BEFORE the call (crash-safe) and outputs after;
                        FLASHINFER_DUMP_INCLUDE / _EXCLUDE are fnmatch filters on the API name, FLASHINFER_DUMP_MAX_COUNT caps
"""
from __future__ import annotations

import fnmatch
import functools
import inspect
import json
import os
import sys
import threading
import time
from typing import Any, Callable, Dict, List, Optional

import torch

_LEVEL = int(os.environ.get("FLASHINFER_LOGLEVEL", "0") or 0)
_DEST = os.environ.get("FLASHINFER_LOGDEST", "stdout")
_DUMP_DIR = os.environ.get("FLASHINFER_DUMP_DIR")
_DUMP_MAX = int(os.environ.get("FLASHINFER_DUMP_MAX_COUNT", "1000"))
_DUMP_INC = os.environ.get("FLASHINFER_DUMP_INCLUDE", "*")
_DUMP_EXC = os.environ.get("FLASHINFER_DUMP_EXCLUDE", "")
_lock = threading.Lock()
_counter = 0
_stream = None
_REGISTRY: Dict[str, Callable] = {}


def _out():
    global _stream
    if _stream is None:
        if _DEST == "stdout":
            _stream = sys.stdout
        elif _DEST == "stderr":
            _stream = sys.stderr
        else:
            _stream = open(_DEST.replace("%i", str(os.getpid())), "a")
    return _stream


def _log(msg: str) -> None:
    with _lock:
        s = _out()
        s.write(f"[flashinfer_b200 {time.strftime('%H:%M:%S')}] {msg}\n")
        s.flush()


def _capturing() -> bool:
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


def tensor_summary(t: torch.Tensor, level: int) -> str:
    s = f"Tensor(shape={tuple(t.shape)}, dtype={str(t.dtype).replace('torch.', '')}, device={t.device}, stride={t.stride()}"
    if level >= 5 and t.numel() and not _capturing() and (t.is_floating_point() or t.dtype in (torch.int32, torch.int64)):
        try:
            from .utils import tensor_stats

            st = tensor_stats(t)
            s += f", min={st['min']:.6g}, max={st['max']:.6g}, mean={st['mean']:.6g}, nan={st['nan']}, inf={st['inf']}"
        except Exception as e:  # noqa: BLE001
            s += f", stats_unavailable={type(e).__name__}"
    return s + ")"


def _fmt(v: Any, level: int) -> str:
    if isinstance(v, torch.Tensor):
        return tensor_summary(v, level)
    if isinstance(v, (list, tuple)) and v and all(isinstance(x, torch.Tensor) for x in v):
        return "[" + ", ".join(tensor_summary(x, level) for x in v) + "]"
    r = repr(v)
    return r if len(r) < 200 else r[:197] + "..."


def _collect_tensors(bound: Dict[str, Any]) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in bound.items():
        if isinstance(v, torch.Tensor):
            out[k] = v
        elif isinstance(v, (list, tuple)):
            for i, x in enumerate(v):
                if isinstance(x, torch.Tensor):
                    out[f"{k}.{i}"] = x
    return out


def _dump(dirpath: str, tag: str, tensors: Dict[str, torch.Tensor], meta: Dict[str, Any]) -> None:
    os.makedirs(dirpath, exist_ok=True)
    torch.save({k: v.detach().cpu() for k, v in tensors.items()}, os.path.join(dirpath, f"{tag}.pt"))
    with open(os.path.join(dirpath, f"{tag}.json"), "w") as f:
        json.dump(meta, f, indent=1, default=repr)


def flashinfer_api(fn: Optional[Callable] = None, *, name: Optional[str] = None):
    """Decorator for public APIs.  Logs (and optionally dumps) the inputs before the call so that a crashing kernel
    still leaves its arguments on disk."""

    def deco(f: Callable) -> Callable:
        api = name or f"{f.__module__.replace('flashinfer_b200.', '')}.{f.__qualname__}"
        _REGISTRY[api] = f
        if _LEVEL <= 0:
            return f
        sig = None

        @functools.wraps(f)
        def wrapper(*args, **kwargs):
            nonlocal sig
            global _counter
            if _LEVEL == 1:
                _log(api)
                return f(*args, **kwargs)
            if sig is None:
                try:
                    sig = inspect.signature(f)
                except (TypeError, ValueError):
                    sig = False
            bound: Dict[str, Any] = {}
            if sig:
                try:
                    ba = sig.bind(*args, **kwargs)
                    bound = dict(ba.arguments)
                except TypeError:
                    bound = {f"arg{i}": a for i, a in enumerate(args)} | kwargs
            bound.pop("self", None)
            _log(f"{api}(" + ", ".join(f"{k}={_fmt(v, _LEVEL)}" for k, v in bound.items()) + ")")
            dump_dir = None
            deferred = None
            if _DUMP_DIR and _capturing() and fnmatch.fnmatch(api, _DUMP_INC) and not (_DUMP_EXC and fnmatch.fnmatch(api, _DUMP_EXC)):
                # inside torch.cuda.graph(...): no D2H copies may be captured -> keep references, write after replay
                with _lock:
                    _counter += 1
                    idx = _counter
                if idx <= _DUMP_MAX:
                    deferred = {"dir": os.path.join(_DUMP_DIR, f"{idx:06d}_{api.replace('.', '_')}"), "api": api, "index": idx,
                                "inputs": _collect_tensors(bound), "outputs": {},
                                "scalars": {k: v for k, v in bound.items() if isinstance(v, (int, float, str, bool, type(None)))}}
            if _DUMP_DIR and not _capturing() and fnmatch.fnmatch(api, _DUMP_INC) and not (_DUMP_EXC and fnmatch.fnmatch(api, _DUMP_EXC)):
                with _lock:
                    _counter += 1
                    idx = _counter
                if idx <= _DUMP_MAX:
                    dump_dir = os.path.join(_DUMP_DIR, f"{idx:06d}_{api.replace('.', '_')}")
                    meta = {"api": api, "index": idx, "scalars": {k: v for k, v in bound.items() if isinstance(v, (int, float, str, bool, type(None)))}}
                    _dump(dump_dir, "inputs", _collect_tensors(bound), meta)
            out = f(*args, **kwargs)
            if _LEVEL >= 3:
                _log(f"{api} -> {_fmt(out, _LEVEL)}")
            if dump_dir is not None:
                outs = out if isinstance(out, (list, tuple)) else [out]
                _dump(dump_dir, "outputs", {f"out{i}": o for i, o in enumerate(outs) if isinstance(o, torch.Tensor)}, {"api": api})
            if deferred is not None:
                outs = out if isinstance(out, (list, tuple)) else [out]
                deferred["outputs"] = {f"out{i}": o for i, o in enumerate(outs) if isinstance(o, torch.Tensor)}
                with _lock:
                    _PENDING_GRAPH_DUMPS.append(deferred)
            return out

        wrapper.__fib200_api_wrapper__ = True
        return wrapper

    return deco(fn) if fn is not None else deco


_PENDING_GRAPH_DUMPS: List[Dict[str, Any]] = []
_FLUSH_COUNTS: Dict[str, int] = {}


def flush_graph_dumps(synchronize: bool = True) -> int:
    """Write the dumps that were deferred during CUDA-graph capture (reference api_logging.py:1075): after ``g.replay()`` the
    recorded input / output tensors hold that replay's values; they are written to the original dump directory (latest
    flush) and to ``graph_flushes/flush_XXXX/`` under it (history).  Returns the number of API calls written."""
    if synchronize and torch.cuda.is_available():
        torch.cuda.synchronize()
    with _lock:
        pending = list(_PENDING_GRAPH_DUMPS)
    for d in pending:
        n = _FLUSH_COUNTS.get(d["dir"], 0)
        _FLUSH_COUNTS[d["dir"]] = n + 1
        meta = {"api": d["api"], "index": d["index"], "scalars": d["scalars"], "graph_flush": n}
        for target in (d["dir"], os.path.join(d["dir"], "graph_flushes", f"flush_{n:04d}")):
            _dump(target, "inputs", d["inputs"], meta)
            _dump(target, "outputs", d["outputs"], {"api": d["api"]})
    return len(pending)


def clear_graph_dumps() -> int:
    """Forget the deferred graph dumps (the captured tensors themselves belong to PyTorch).  Returns how many were dropped."""
    with _lock:
        n = len(_PENDING_GRAPH_DUMPS)
        _PENDING_GRAPH_DUMPS.clear()
        _FLUSH_COUNTS.clear()
    return n


def replay_from_dump(dump_dir: str, device: str = "cuda", compare: bool = True, rtol: float = 1e-2, atol: float = 1e-2) -> Dict[str, Any]:
    """Re-run one dumped call (functional APIs only) and compare against the dumped outputs."""
    with open(os.path.join(dump_dir, "inputs.json")) as f:
        meta = json.load(f)
    fn = _REGISTRY.get(meta["api"])
    if fn is None:
        raise KeyError(f"API {meta['api']} is not registered in this process (import its module first)")
    tensors = torch.load(os.path.join(dump_dir, "inputs.pt"))
    kwargs: Dict[str, Any] = dict(meta.get("scalars", {}))
    lists: Dict[str, Dict[int, torch.Tensor]] = {}
    for k, v in tensors.items():
        if "." in k:
            base, i = k.rsplit(".", 1)
            lists.setdefault(base, {})[int(i)] = v.to(device)
        else:
            kwargs[k] = v.to(device)
    for base, d in lists.items():
        kwargs[base] = [d[i] for i in sorted(d)]
    out = fn(**kwargs)
    res: Dict[str, Any] = {"api": meta["api"], "output": out}
    ref_path = os.path.join(dump_dir, "outputs.pt")
    if compare and os.path.exists(ref_path):
        ref = torch.load(ref_path)
        outs = out if isinstance(out, (list, tuple)) else [out]
        ok = True
        for i, o in enumerate(outs):
            r = ref.get(f"out{i}")
            if r is not None and isinstance(o, torch.Tensor):
                ok = ok and torch.allclose(o.detach().float().cpu(), r.float(), rtol=rtol, atol=atol, equal_nan=True)
        res["match"] = ok
    return res


def replay_sequence(root: str, device: str = "cuda", **kw) -> List[Dict[str, Any]]:
    return [replay_from_dump(os.path.join(root, d), device, **kw) for d in sorted(os.listdir(root))
            if os.path.isdir(os.path.join(root, d))]


# ------------------------------------------------------------------------------------------------------------------
# Package-wide instrumentation.  The reference puts ``@flashinfer_api`` on every public function and on the wrappers'
# plan / run methods (flashinfer/api_logging.py:2364-2455 lists what replay can re-create).  Here the same set is decorated
# in one sweep at package import: with FLASHINFER_LOGLEVEL=0 ``flashinfer_api`` hands the function back unchanged (zero
# overhead) and only fills the replay registry; with a level > 0 - at import or later through ``set_level`` - the public
# names are rebound to the logging wrappers in their home module and wherever the package re-exports them.
# ------------------------------------------------------------------------------------------------------------------
PUBLIC_API_MODULES = (
    "activation", "cascade", "concat_ops", "decode", "deep_gemm", "dsv3_ops", "gdn", "norm", "page", "pod", "prefill", "rope",
    "sampling", "sparse", "topk", "xqa", "attention._core", "mla._core", "gemm.dense", "gemm.lowp", "gemm.grouped",
    "gemm.decode_linear", "fused_moe.core", "quantization.fp4", "quantization.fp8", "quantization.packbits",
    "mamba.selective_state_update", "mamba.ssd_combined", "comm.allreduce", "comm.trtllm_ar", "comm.vllm_ar", "comm.cuda_ipc", "comm.dlpack_utils", "comm.trtllm_mnnvl_ar", "comm.alltoall", "comm.collectives", "comm.dcp_alltoall", "comm.mixed_comm",
    "comm.gemm_allreduce", "logits_processor.pipeline",
)
_WRAPPER_METHODS = ("plan", "run", "forward", "begin_forward", "dispatch", "combine")
_ORIGINALS: Dict[str, Any] = {}     # api name -> (owner object, attribute, undecorated callable)
_PKG = __name__.rsplit(".", 1)[0]


def _rebind_everywhere(old, new) -> None:
    for name, mod in list(sys.modules.items()):
        if mod is None or not (name == _PKG or name.startswith(_PKG + ".")):
            continue
        for attr, val in list(vars(mod).items()):
            if val is old:
                setattr(mod, attr, new)


def _unwrap(f):
    """Undo OUR logging wrapper only (functools.wraps also leaves ``__wrapped__`` on contextmanagers, lru_caches, ...)."""
    return f.__wrapped__ if getattr(f, "__fib200_api_wrapper__", False) else f


def instrument() -> int:
    """Decorate every public op / wrapper method of the package (idempotent).  Returns the number of instrumented callables."""
    import importlib

    n = 0
    for short in PUBLIC_API_MODULES:
        try:
            mod = importlib.import_module(f"{_PKG}.{short}")
        except Exception:  # noqa: BLE001  (optional sub-module)
            continue
        for attr, obj in list(vars(mod).items()):
            if attr.startswith("_"):
                continue
            if inspect.isfunction(obj) and getattr(obj, "__module__", None) == mod.__name__:
                api = f"{short}.{attr}"
                base = _ORIGINALS.setdefault(api, (mod, attr, _unwrap(obj)))[2]
                new = flashinfer_api(base, name=api)
                if new is not obj:
                    _rebind_everywhere(obj, new)
                n += 1
            elif inspect.isclass(obj) and getattr(obj, "__module__", None) == mod.__name__:
                for meth in _WRAPPER_METHODS:
                    f = obj.__dict__.get(meth)
                    if not inspect.isfunction(f):
                        continue
                    api = f"{short}.{attr}.{meth}"
                    base = _ORIGINALS.setdefault(api, (obj, meth, _unwrap(f)))[2]
                    new = flashinfer_api(base, name=api)
                    if new is not f:
                        setattr(obj, meth, new)
                    n += 1
    return n


def set_level(level: int, dest: Optional[str] = None, dump_dir: Optional[str] = None) -> int:
    """Change the log level at run time (the environment variables are only the defaults) and re-instrument the package."""
    global _LEVEL, _DEST, _DUMP_DIR, _stream
    _LEVEL = int(level)
    if dest is not None:
        _DEST, _stream = dest, None
    if dump_dir is not None:
        _DUMP_DIR = dump_dir or None
    # put the undecorated callables back first, then decorate at the new level
    for api, (owner, attr, base) in _ORIGINALS.items():
        cur = getattr(owner, attr, None)
        if cur is not base and cur is not None:
            if inspect.isclass(owner):
                setattr(owner, attr, base)
            else:
                _rebind_everywhere(cur, base)
    return instrument()


def registered_apis() -> List[str]:
    return sorted(_REGISTRY)
