"""Tactic auto-tuner.  Parity: reference flashinfer/autotuner.py (TunableRunner / TuningConfig / AutoTuner.choose_one
:1045 / autotune() :465, JSON save/load of tuned configs).

B200-first differences: tactics are plain hashable values (e.g. the N-tile width of a tcgen05 GEMM), profiling uses
CUDA events with a cold L2 and optional CUDA-graph capture, and the cache key buckets dynamic dims to powers of two so a
serving engine tunes once per bucket.  Outside ``with autotune():`` the tuner only *looks up* cached choices (falls back
to the runner's heuristic tactic ``-1``), so the hot path never profiles.
"""
from __future__ import annotations

import contextlib
import json
import os
import threading
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch


def next_positive_power_of_2(x: int) -> int:
    return 1 if x < 1 else 1 << (int(x) - 1).bit_length()


def last_positive_power_of_2(x: int) -> int:
    n = next_positive_power_of_2(x)
    return n if n == x else n // 2


def get_power_of_2_num_tokens_buckets(max_num_tokens: int) -> Tuple[int, ...]:
    out, m = [], next_positive_power_of_2(max_num_tokens)
    while m >= 1:
        out.append(m)
        m //= 2
    return tuple(out)


@dataclass(frozen=True)
class DynamicTensorSpec:
    """``input_idx[i]`` / ``dim_idx[i]`` name tensor dims that vary at run time; ``gen_tuning_buckets`` lists the
    sizes to profile and ``map_to_tuning_buckets`` maps a live size to its bucket."""
    input_idx: Tuple[int, ...]
    dim_idx: Tuple[int, ...]
    gen_tuning_buckets: Tuple[int, ...] = ()
    map_to_tuning_buckets: Callable[[int], int] = next_positive_power_of_2
    tensor_initializers: Tuple[Callable, ...] = ()


@dataclass(frozen=True)
class ConstraintSpec:
    input_idx: int
    dim_idx: int
    infer_shape: Callable[[List[Tuple[int, ...]]], int]


@dataclass
class TuningConfig:
    dynamic_tensor_specs: Tuple[DynamicTensorSpec, ...] = ()
    constraint_specs: Tuple[ConstraintSpec, ...] = ()
    use_cuda_graph: bool = False
    use_cold_l2_cache: bool = True
    synthesize_buckets: bool = True   # False: only ever profile live inputs (runners that hold state sized by a dynamic dim)


class TunableRunner:
    """A kernel family with selectable tactics.  ``tactic == -1`` must always work (the built-in heuristic)."""

    def get_valid_tactics(self, inputs: List[torch.Tensor], profile: "OptimizationProfile") -> List[Any]:
        return [-1]

    def forward(self, inputs: List[torch.Tensor], tactic: Any = -1, do_preparation: bool = False, **kwargs):
        raise NotImplementedError

    def get_cache_key_extras(self, inputs: List[torch.Tensor]) -> Tuple:
        """Extra hashable state that distinguishes cached tactics beyond the input shapes (default: none)."""
        return ()

    def __call__(self, inputs, **kwargs):
        return self.forward(inputs, **kwargs)

    def __hash__(self):
        return hash(type(self).__name__)


@dataclass
class StaticDim:
    """A dimension with one value (reference autotuner.py:332)."""
    val: int

    def _opt(self) -> int:
        return self.val


@dataclass(unsafe_hash=True)
class DynamicDim:
    """Range of one dimension (reference autotuner.py:340)."""
    min: int
    opt: int
    max: int

    def _opt(self) -> int:
        return self.opt


@dataclass
class FakeTensor:
    """Shape-only stand-in used when profiles are enumerated without allocating (reference autotuner.py:375)."""
    dtype: torch.dtype
    device: torch.device
    shape: List[Any]


@dataclass
class AutoTunerStatistics:
    """Counters the tuner keeps (reference autotuner.py:697)."""
    cache_misses: int = 0
    cache_miss_config_collection: Dict[str, set] = field(default_factory=dict)
    failed_profiling_count: Dict[str, set] = field(default_factory=dict)
    tuned_op_total_configs: Dict[str, int] = field(default_factory=dict)
    tuned_op_successful_configs: Dict[str, int] = field(default_factory=dict)

    def __str__(self) -> str:
        return (f"Cache misses: {self.cache_misses}\nTuned ops: {dict(self.tuned_op_total_configs)}\n"
                f"Successful: {dict(self.tuned_op_successful_configs)}\n")


@dataclass
class OptimizationProfile:
    shapes: List[Tuple[int, ...]] = field(default_factory=list)
    tensor_initializers: List[Any] = field(default_factory=list)

    def key(self) -> Tuple:
        return tuple(self.shapes)

    def get_opt_shapes(self) -> Tuple:
        """Shapes with every Dim resolved to its tuning value (plain ints pass through)."""
        return tuple(tuple(d._opt() if hasattr(d, "_opt") else int(d) for d in shp) for shp in self.shapes)

    def get_hash_key(self) -> Tuple:
        return self.get_opt_shapes()


_measure = threading.local()


def is_in_profile_measurement() -> bool:
    """True while the tuner is timing a candidate on this thread (ops use it to skip logging / validation side effects)."""
    return getattr(_measure, "depth", 0) > 0


@contextlib.contextmanager
def _profile_measurement_scope():
    _measure.depth = getattr(_measure, "depth", 0) + 1
    try:
        yield
    finally:
        _measure.depth -= 1


def _collect_metadata() -> Dict[str, str]:
    """What a tuned-config file was measured on; a mismatch at load time is reported, not fatal."""
    from .version import __version__

    meta = {"package": f"flashinfer_b200 {__version__}", "torch": torch.__version__, "cuda": str(torch.version.cuda)}
    if torch.cuda.is_available():
        prop = torch.cuda.get_device_properties(torch.cuda.current_device())
        meta.update(device=prop.name, sm=f"{prop.major}{prop.minor}", sm_count=str(prop.multi_processor_count))
    else:
        meta["device"] = "cpu"
    return meta


def get_config_path(is_module: bool = False) -> str:
    """Where shipped tuned configs for the current device live: ``tuning_configs/<device>.json`` next to this file
    (``is_module`` returns the dotted module-style name the reference uses for the same file)."""
    dev = torch.cuda.get_device_name().replace(" ", "_") if torch.cuda.is_available() else "cpu"
    name = f"v0_1_{dev}"
    return f"flashinfer_b200.tuning_configs.{name}" if is_module else os.path.join(os.path.dirname(__file__), "tuning_configs", name + ".json")


def _tactic_to_json(t):
    return {"__tuple__": [_tactic_to_json(x) for x in t]} if isinstance(t, tuple) else t


def _json_to_tactic(v):
    if isinstance(v, dict) and "__tuple__" in v:
        return tuple(_json_to_tactic(x) for x in v["__tuple__"])
    return tuple(_json_to_tactic(x) for x in v) if isinstance(v, list) else v


class AutoTuner:
    _instance: Optional["AutoTuner"] = None
    _lock = threading.Lock()

    def __init__(self, warmup: int = 3, repeat: int = 10, stream_delay_micro_secs: int = 0) -> None:
        self.is_tuning_mode = False
        self.warmup, self.repeat = warmup, repeat
        self.profiling_cache: Dict[Tuple, Tuple[int, Any, float]] = {}
        self.stats = {"hits": 0, "misses": 0, "profiled": 0, "failed": 0}
        self._flush = None
        self._overrides = threading.local()

    def reset_statistics(self) -> None:
        self.stats = {"hits": 0, "misses": 0, "profiled": 0, "failed": 0}

    # ---- per-thread bucket overrides pushed by ``autotune(tuning_buckets=..., round_up=...)``
    def _get_override_stack(self) -> List:
        if not hasattr(self._overrides, "stack"):
            self._overrides.stack = []
        return self._overrides.stack

    def _override_tuning_buckets(self) -> Optional[Tuple[int, ...]]:
        for buckets, _ in reversed(self._get_override_stack()):
            if buckets is not None:
                return buckets
        return None

    def _override_round_up(self) -> bool:
        for _, up in reversed(self._get_override_stack()):
            if up is not None:
                return up
        return False

    def get_effective_map_to_tuning_buckets(self, spec=None) -> Callable[[int], int]:
        """The live-size -> bucket map in force: the innermost ``autotune()`` override, else the spec's own mapper, else
        next-power-of-two."""
        buckets, up = self._override_tuning_buckets(), self._override_round_up()
        if buckets is not None:
            from .fused_moe.utils import make_bucket_mapper

            return make_bucket_mapper(buckets, round_map=up)
        gen = tuple(sorted(set(getattr(spec, "gen_tuning_buckets", ()) or ())))
        if up and gen:
            from .fused_moe.utils import make_bucket_mapper

            return make_bucket_mapper(gen, round_map=True)
        fn = getattr(spec, "map_to_tuning_buckets", None)
        return fn if fn is not None else next_positive_power_of_2

    @classmethod
    def get(cls) -> "AutoTuner":
        with cls._lock:
            if cls._instance is None:
                cls._instance = AutoTuner()
                cls._instance._load_shipped()
            return cls._instance

    def _load_shipped(self) -> None:
        """Seed the cache with the tuned-config file shipped for this device (``tuning_configs/``), then with the user's
        ``$FLASHINFER_AUTOTUNER_CACHE``; a missing or unreadable file is not an error."""
        for path in (get_config_path(), os.environ.get("FLASHINFER_AUTOTUNER_CACHE")):
            if path and os.path.exists(path):
                try:
                    self.load_configs(path)
                except (OSError, ValueError, KeyError):
                    pass

    # ---- cache
    def _bucket_shapes(self, inputs: Sequence[Any], cfg: TuningConfig) -> Tuple:
        shapes = [list(t.shape) if isinstance(t, (torch.Tensor, FakeTensor)) else [] for t in inputs]
        for spec in cfg.dynamic_tensor_specs:
            mapper = self.get_effective_map_to_tuning_buckets(spec)
            for ii, dd in zip(spec.input_idx, spec.dim_idx):
                if ii < len(shapes) and dd < len(shapes[ii]):
                    shapes[ii][dd] = mapper(shapes[ii][dd])
        for c in cfg.constraint_specs:
            if c.input_idx < len(shapes) and c.dim_idx < len(shapes[c.input_idx]):
                shapes[c.input_idx][c.dim_idx] = c.infer_shape([tuple(s) for s in shapes])
        return tuple(tuple(s) for s in shapes)

    def _key(self, op: str, runner: TunableRunner, shapes: Tuple, extras: Tuple = ()) -> Tuple:
        return (op, type(runner).__name__, shapes, extras)

    def search_cache(self, op: str, runners: Sequence[TunableRunner], shapes: Tuple, extras: Tuple = ()):
        for i, r in enumerate(runners):
            hit = self.profiling_cache.get(self._key(op, r, shapes, extras))
            if hit is not None:
                return True, i, hit[1]
        return False, 0, -1

    def clear_cache(self) -> None:
        self.profiling_cache.clear()

    def save_configs(self, path: str) -> None:
        """Merge this process' choices into ``path`` (entries already in the file for other shapes are kept) and replace
        the file atomically, so concurrent writers lose at most their own update, never the file."""
        merged: Dict[Tuple, Tuple[int, Any, float]] = {}
        if os.path.exists(path):
            try:
                other = AutoTuner()
                other.load_configs(path)
                merged.update(other.profiling_cache)
            except (OSError, ValueError, KeyError):
                pass
        merged.update(self.profiling_cache)
        rows = [{"op": k[0], "runner": k[1], "shapes": [list(s) for s in k[2]], "extras": _tactic_to_json(tuple(k[3])),
                 "runner_id": v[0], "tactic": _tactic_to_json(v[1]), "ms": v[2]} for k, v in sorted(merged.items(), key=repr)]
        tmp = f"{path}.{os.getpid()}.tmp"
        with open(tmp, "w") as f:
            json.dump({"metadata": _collect_metadata(), "configs": rows}, f, indent=1)
        os.replace(tmp, path)

    def load_configs(self, path: str) -> int:
        with open(path) as f:
            data = json.load(f)
        meta = data.get("metadata", {})
        here = _collect_metadata()
        for field_ in ("device", "sm"):
            if field_ in meta and field_ in here and meta[field_] != here[field_]:
                import warnings

                warnings.warn(f"tuned configs in {path} were measured on {field_}={meta[field_]!r}, this process runs on "
                              f"{here[field_]!r}; the choices are still valid tactics but may not be the fastest", RuntimeWarning)
                break
        for r in data.get("configs", []):
            extras = _json_to_tactic(r.get("extras", []))
            key = (r["op"], r["runner"], tuple(tuple(s) for s in r["shapes"]), tuple(extras) if isinstance(extras, tuple) else ())
            self.profiling_cache[key] = (r.get("runner_id", 0), _json_to_tactic(r["tactic"]), r.get("ms", 0.0))
        return len(data.get("configs", []))

    # ---- profiling
    def _time(self, runner: TunableRunner, inputs, tactic, cfg: TuningConfig, **kwargs) -> float:
        with _profile_measurement_scope():
            return self._time_inner(runner, inputs, tactic, cfg, **kwargs)

    def _time_inner(self, runner: TunableRunner, inputs, tactic, cfg: TuningConfig, **kwargs) -> float:
        if not torch.cuda.is_available():
            import time

            t0 = time.perf_counter()
            runner.forward(inputs, tactic=tactic, **kwargs)
            return (time.perf_counter() - t0) * 1e3
        for _ in range(self.warmup):
            runner.forward(inputs, tactic=tactic, **kwargs)
        torch.cuda.synchronize()
        if cfg.use_cold_l2_cache and self._flush is None:
            self._flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        graph = None
        if cfg.use_cuda_graph:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                runner.forward(inputs, tactic=tactic, **kwargs)
        ts = []
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(self.repeat):
            if cfg.use_cold_l2_cache:
                self._flush.zero_()
            s.record()
            if graph is not None:
                graph.replay()
            else:
                runner.forward(inputs, tactic=tactic, **kwargs)
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        ts.sort()
        return ts[len(ts) // 2]

    def _resized(self, inputs: List[Any], cfg: TuningConfig, spec: DynamicTensorSpec, size: int) -> List[Any]:
        """Inputs with every dimension named by ``spec`` set to ``size`` (fresh tensors from the spec's initialisers,
        default: random normal cast to the dtype) and the constraint dims re-derived."""
        out = list(inputs)
        for n, (ii, dd) in enumerate(zip(spec.input_idx, spec.dim_idx)):
            t = out[ii]
            shape = list(t.shape)
            shape[dd] = size
            init = spec.tensor_initializers[n] if n < len(spec.tensor_initializers) else None
            out[ii] = init(shape, t.dtype, t.device) if init is not None else _default_init(shape, t.dtype, t.device)
        for c in cfg.constraint_specs:
            t = out[c.input_idx]
            want = c.infer_shape([tuple(x.shape) if isinstance(x, torch.Tensor) else () for x in out])
            if t.shape[c.dim_idx] != want:
                shape = list(t.shape)
                shape[c.dim_idx] = want
                out[c.input_idx] = _default_init(shape, t.dtype, t.device)
        return out

    def _profile(self, custom_op, runners, cfg, inputs, shapes, extras, **kwargs):
        best = (float("inf"), 0, -1)
        prof = OptimizationProfile([tuple(s) for s in shapes])
        for rid, r in enumerate(runners):
            for tac in r.get_valid_tactics(inputs, prof):
                try:
                    ms = self._time(r, inputs, tac, cfg, **kwargs)
                    self.stats["profiled"] += 1
                except Exception:  # noqa: BLE001 - a tactic that cannot run is simply not a candidate
                    self.stats["failed"] += 1
                    continue
                if ms < best[0]:
                    best = (ms, rid, tac)
        ms, rid, tac = best
        self.profiling_cache[self._key(custom_op, runners[rid], shapes, extras)] = (rid, tac, ms)
        return rid, tac

    def choose_one(self, custom_op: str, runners: Sequence[TunableRunner], tuning_config: TuningConfig,
                   inputs: List[torch.Tensor], extras: Tuple = (), **kwargs) -> Tuple[TunableRunner, Any]:
        """Returns ``(runner, tactic)``: the cached best, or - inside ``autotune()`` - the freshly profiled best.

        In tuning mode a miss profiles the live shape's bucket and, when the config names ``gen_tuning_buckets`` (or an
        ``autotune(tuning_buckets=...)`` override is active), every other listed bucket as well on synthetic inputs, so
        one warm-up pass covers the whole serving range (reference autotuner.py:1045-1330)."""
        shapes = self._bucket_shapes(inputs, tuning_config)
        hit, rid, tactic = self.search_cache(custom_op, runners, shapes, extras)
        if hit:
            self.stats["hits"] += 1
            return runners[rid], tactic
        if not self.is_tuning_mode:
            self.stats["misses"] += 1
            return runners[0], -1
        rid, tactic = self._profile(custom_op, runners, tuning_config, inputs, shapes, extras, **kwargs)
        for spec in (tuning_config.dynamic_tensor_specs if tuning_config.synthesize_buckets else ()):
            sizes = self._override_tuning_buckets() or tuple(spec.gen_tuning_buckets or ())
            for size in sizes:
                try:
                    alt = self._resized(inputs, tuning_config, spec, int(size))
                except Exception:  # noqa: BLE001 - inputs that cannot be synthesised are tuned when they show up live
                    continue
                alt_shapes = self._bucket_shapes(alt, tuning_config)
                if not self.search_cache(custom_op, runners, alt_shapes, extras)[0]:
                    self._profile(custom_op, runners, tuning_config, alt, alt_shapes, extras, **kwargs)
        return runners[rid], tactic


def _default_init(shape, dtype, device):
    if dtype.is_floating_point and dtype.itemsize >= 2:
        return torch.randn(shape, device=device, dtype=torch.float32).to(dtype)
    if dtype.is_floating_point:                      # fp8: cast from a bounded normal
        return (torch.randn(shape, device=device) * 0.5).to(dtype)
    return torch.zeros(shape, dtype=dtype, device=device)


@contextlib.contextmanager
def autotune(tune_mode: bool = True, cache: Optional[str] = None, tuning_buckets: Optional[Sequence[int]] = None,
             round_up: Optional[bool] = None, cache_path: Optional[str] = None):
    """``with autotune():`` profiles every tunable op reached inside the block and caches the winners.

    ``cache`` (or ``$FLASHINFER_AUTOTUNER_CACHE``; ``cache_path`` is the older spelling) names a JSON file that is loaded on
    entry and merged + saved on exit when tuning.  ``tuning_buckets`` replaces every op's dynamic-dimension buckets inside
    the block; ``round_up`` picks ceil instead of floor when mapping a live size onto them.  Overrides nest per thread and
    ``None`` inherits from the enclosing block (reference autotuner.py:465-660)."""
    if tuning_buckets is not None:
        tuning_buckets = tuple(sorted({int(b) for b in tuning_buckets}))
        if not tuning_buckets:
            raise ValueError("tuning_buckets must contain at least one value")
    tuner = AutoTuner.get()
    path = cache or cache_path or os.environ.get("FLASHINFER_AUTOTUNER_CACHE")
    if path and os.path.exists(path):
        tuner.load_configs(path)
    old = tuner.is_tuning_mode
    tuner.is_tuning_mode = tune_mode
    stack = tuner._get_override_stack()
    stack.append((tuning_buckets, round_up))
    try:
        yield tuner
    finally:
        stack.pop()
        tuner.is_tuning_mode = old
        if path and tune_mode:
            tuner.save_configs(path)


def load_from_file(key) -> Tuple[bool, int, Any]:
    """Look ``key = (op, runner class name, shapes, extras)`` up in the shipped config file for this device."""
    path = get_config_path()
    if not os.path.exists(path):
        return False, 0, -1
    t = AutoTuner()
    t.load_configs(path)
    hit = t.profiling_cache.get(tuple(key))
    return (True, hit[0], hit[1]) if hit is not None else (False, 0, -1)
