"""Native module build + load layer (sm_100a only).

Design (B200-first, replaces the reference's JitSpec/ninja/TVM-FFI stack,
cf. reference flashinfer/jit/core.py:217-402 and flashinfer/jit/cpp_ext.py:238-343):

* One target arch (``sm_100a``) => no arch matrix, no backend zoo.  Every native
  module is a small set of ``.cu``/``.cpp`` sources compiled by ``nvcc`` straight
  into an in-tree shared object ``flashinfer_b200/_lib/<name>.so``.
* Kernels export a plain C ABI (``extern "C" int fn(void*, int64_t, double, ...)``)
  with a *uniform* calling convention: pointers are ``void*``, all integers are
  ``int64_t``, all floats are ``double``, the last argument is the CUDA stream.
  That lets one generic ctypes caller marshal any op (no per-op binding code),
  and keeps compile times to seconds because no torch headers are included.
* Staleness is tracked with a content hash (sources + headers + flags) stored
  next to the ``.so``; ``build_all()`` is the AOT entry point used by
  ``__graft_entry__.build()``, ``load()`` JIT-builds on first use when ``nvcc``
  is available and the ``.so`` is stale/missing (FileLock-protected, like the
  reference's JIT cache lock).
* On a GPU box a missing module is a hard error (no silent eager fallback).

Package layout (reference flashinfer/jit/): this file is the build + load machinery; ``core`` holds the reference's JitSpec names
on top of it, ``env`` the directory constants, ``cpp_ext`` the toolchain probes and the ninja writer, ``cubin_loader`` the
(empty, nothing is downloaded) artifact hooks, ``attention`` the user-variant generator, ``utils`` the dtype tables.
"""
from __future__ import annotations

import contextlib
import ctypes
import hashlib
import os
import shutil
import subprocess
import threading
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import torch

_PKG = Path(__file__).resolve().parent.parent      # the flashinfer_b200 package directory
CSRC = _PKG / "csrc"
LIB_DIR = _PKG / "_lib"
INCLUDE_DIRS = [CSRC / "include"]

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON_NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-lineinfo",
    "--expt-relaxed-constexpr",
    "--expt-extended-lambda",
    "-Xcompiler",
    "-fPIC",
    "-Xcompiler",
    "-Wno-unused-function",
    "--threads",
    "2",
    "-diag-suppress",
    "177,550",
]


def debug_flags() -> List[str]:
    """``FIB200_JIT_DEBUG=1``: device debug build (``-G -g --ptxas-options=-v``), like the reference's FLASHINFER_JIT_DEBUG; the
    flags enter the content hash, so switching the variable rebuilds the affected modules."""
    return ["-G", "-g", "--ptxas-options=-v"] if os.environ.get("FIB200_JIT_DEBUG", "0") == "1" else []


class MissingNativeModuleError(RuntimeError):
    """Raised when a native module is required but is neither built nor buildable."""


@dataclass
class ModuleSpec:
    name: str
    sources: Sequence[str]  # relative to csrc/
    extra_flags: Sequence[str] = field(default_factory=list)
    ldflags: Sequence[str] = field(default_factory=list)
    deps: Sequence[str] = field(default_factory=list)  # files #included by the sources (hashed, not compiled)

    @property
    def so_path(self) -> Path:
        return LIB_DIR / f"{self.name}.so"

    @property
    def hash_path(self) -> Path:
        return LIB_DIR / f"{self.name}.hash"

    def source_paths(self) -> List[Path]:
        return [CSRC / s for s in self.sources]

    def content_hash(self) -> str:
        h = hashlib.sha256()
        for p in self.source_paths() + [CSRC / d for d in self.deps]:
            h.update(p.name.encode())
            h.update(p.read_bytes())
        for inc in INCLUDE_DIRS:
            for p in sorted(inc.rglob("*")):
                if p.is_file():
                    h.update(p.name.encode())
                    h.update(p.read_bytes())
        h.update(" ".join(list(ARCH_FLAGS) + list(COMMON_NVCC_FLAGS) + debug_flags() + list(self.extra_flags) + list(self.ldflags)).encode())
        return h.hexdigest()

    def is_fresh(self) -> bool:
        if not self.so_path.exists() or not self.hash_path.exists():
            return False
        try:
            return self.hash_path.read_text().strip() == self.content_hash()
        except OSError:
            return False

    def nvcc_command(self) -> List[str]:
        nvcc = os.environ.get("FIB200_NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        cmd = [nvcc, "-shared", *ARCH_FLAGS, *COMMON_NVCC_FLAGS, *debug_flags()]
        for inc in INCLUDE_DIRS:
            cmd += ["-I", str(inc)]
        cmd += list(self.extra_flags)
        cmd += [str(p) for p in self.source_paths()]
        cmd += ["-o", str(self.so_path)]
        cmd += list(self.ldflags)
        return cmd


# ----------------------------------------------------------------------------
# Module registry.  Every native module of the framework is declared here so
# that `build_all()` (AOT) and the CLI `module-status` can enumerate them
# (reference: JitSpecRegistry, flashinfer/jit/core.py).
# ----------------------------------------------------------------------------
REGISTRY: Dict[str, ModuleSpec] = {}


def register(spec: ModuleSpec) -> ModuleSpec:
    REGISTRY[spec.name] = spec
    return spec


register(ModuleSpec("norm", ["elementwise/norm.cu"]))
register(ModuleSpec("activation", ["elementwise/activation.cu"]))
register(ModuleSpec("rope", ["elementwise/rope.cu"]))
register(ModuleSpec("page", ["elementwise/page.cu"]))
register(ModuleSpec("cascade", ["elementwise/cascade.cu"]))
register(ModuleSpec("quantization", ["elementwise/quantization.cu"]))
register(ModuleSpec("sampling", ["elementwise/sampling.cu"]))
register(ModuleSpec("topk", ["elementwise/topk.cu"]))
register(ModuleSpec("planner", ["runtime/planner.cpp"]))
register(ModuleSpec("runtime", ["runtime/runtime.cu"]))
register(ModuleSpec("gemm_sm100", ["gemm/gemm_bf16_sm100.cu"]))
register(ModuleSpec("decode_linear_sm100", ["gemm/decode_linear_sm100.cu"]))
register(ModuleSpec("decode_linear_sm100_prof", ["gemm/decode_linear_sm100.cu"], extra_flags=["-DFIB200_ENABLE_PROFILER"]))  # intra-kernel profiler build
register(ModuleSpec("decode_sm100", ["attention/decode_sm100.cu"]))
register(ModuleSpec("prefill_sm100", ["attention/prefill_sm100.cu"]))
register(ModuleSpec("mla_sm100", ["attention/mla_sm100.cu"]))
register(ModuleSpec("pod_sm100", ["attention/pod_sm100.cu"], deps=["attention/prefill_sm100.cu", "attention/decode_sm100.cu"]))
register(ModuleSpec("gemm_blockscaled_sm100", ["gemm/gemm_blockscaled_sm100.cu"]))
register(ModuleSpec("grouped_gemm_sm100", ["gemm/grouped_gemm_sm100.cu"]))
register(ModuleSpec("moe", ["moe/routing.cu"]))
register(ModuleSpec("ssm", ["elementwise/ssm.cu"]))
register(ModuleSpec("attention_generic", ["attention/generic_attention.cu"]))
register(ModuleSpec("comm_allreduce", ["comm/allreduce.cu"]))
register(ModuleSpec("comm_alltoall", ["comm/moe_a2a.cu"]))
register(ModuleSpec("comm_collectives", ["comm/collectives.cu"]))
register(ModuleSpec("gemm_comm_sm100", ["gemm/gemm_allreduce_sm100.cu"]))
register(ModuleSpec("gemm_allgather_sm100", ["gemm/gemm_allgather_sm100.cu"]))


def _existing(spec: ModuleSpec) -> bool:
    return all(p.exists() for p in spec.source_paths())


_build_lock = threading.Lock()


def have_nvcc() -> bool:
    return bool(os.environ.get("FIB200_NVCC") or shutil.which("nvcc") or Path("/usr/local/cuda/bin/nvcc").exists())


def build_module(spec: ModuleSpec, verbose: bool = False, force: bool = False) -> Path:
    """Compile one module in-tree (no-op when the content hash matches)."""
    from filelock import FileLock

    LIB_DIR.mkdir(exist_ok=True)
    if not force and spec.is_fresh():
        return spec.so_path
    if not have_nvcc():
        raise MissingNativeModuleError(f"native module '{spec.name}' is stale/missing and nvcc is unavailable")
    with FileLock(str(LIB_DIR / f"{spec.name}.lock")):
        if not force and spec.is_fresh():
            return spec.so_path
        cmd = spec.nvcc_command()
        if verbose:
            print("[fib200.jit]", " ".join(cmd), flush=True)
        tmp_out = spec.so_path.with_suffix(".so.tmp")
        cmd[cmd.index("-o") + 1] = str(tmp_out)
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed for module '{spec.name}':\n{proc.stdout}\n{proc.stderr}")
        if verbose and proc.stderr.strip():
            print(proc.stderr)
        os.replace(tmp_out, spec.so_path)
        spec.hash_path.write_text(spec.content_hash())
    return spec.so_path


def build_all(verbose: bool = False, jobs: Optional[int] = None, force: bool = False) -> List[str]:
    """AOT-build every registered module whose sources exist. Returns built names."""
    specs = [s for s in REGISTRY.values() if _existing(s)]
    jobs = jobs or max(1, min(len(specs), (os.cpu_count() or 4) // 2))
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        list(ex.map(lambda s: build_module(s, verbose=verbose, force=force), specs))
    return [s.name for s in specs]


def module_status() -> Dict[str, str]:
    out = {}
    for name, spec in REGISTRY.items():
        if not _existing(spec):
            out[name] = "no-source"
        elif spec.is_fresh():
            out[name] = "built"
        elif spec.so_path.exists():
            out[name] = "stale"
        else:
            out[name] = "missing"
    return out


# ----------------------------------------------------------------------------
# ctypes loader with a uniform calling convention.
# ----------------------------------------------------------------------------
class NativeModule:
    """Thin wrapper over ``ctypes.CDLL`` implementing the uniform C ABI."""

    def __init__(self, name: str, path: Path):
        self.name = name
        self.path = path
        self._dll = ctypes.CDLL(str(path))
        self._fns: Dict[str, ctypes._CFuncPtr] = {}
        try:
            self._last_error = self._dll.fib200_last_error
            self._last_error.restype = ctypes.c_char_p
        except AttributeError:
            self._last_error = None

    def fn(self, sym: str):
        f = self._fns.get(sym)
        if f is None:
            f = getattr(self._dll, sym)
            f.restype = ctypes.c_int
            self._fns[sym] = f
        return f

    def has(self, sym: str) -> bool:
        try:
            self.fn(sym)
            return True
        except AttributeError:
            return False

    def call(self, sym: str, *args) -> None:
        """Marshal ``args`` (Tensor/None -> void*, int -> int64, float -> double) and
        raise on a non-zero return code."""
        cargs = []
        for a in args:
            if a is None:
                cargs.append(ctypes.c_void_p(0))
            elif isinstance(a, torch.Tensor):
                cargs.append(ctypes.c_void_p(a.data_ptr()))
            elif isinstance(a, bool):
                cargs.append(ctypes.c_int64(int(a)))
            elif isinstance(a, int):
                cargs.append(ctypes.c_int64(a))
            elif isinstance(a, float):
                cargs.append(ctypes.c_double(a))
            elif isinstance(a, (ctypes.c_void_p, ctypes.c_int64, ctypes.c_double)):
                cargs.append(a)
            elif hasattr(a, "v") and type(a).__name__ == "_ptr":  # raw device address
                cargs.append(ctypes.c_void_p(a.v))
            else:
                raise TypeError(f"cannot marshal argument of type {type(a)} for {self.name}.{sym}")
        rc = self.fn(sym)(*cargs)
        if rc != 0:
            msg = ""
            if self._last_error is not None:
                raw = self._last_error()
                msg = raw.decode() if raw else ""
            raise RuntimeError(f"{self.name}.{sym} failed (code {rc}): {msg}")


_loaded: Dict[str, NativeModule] = {}


_redirect = threading.local()


@contextlib.contextmanager
def redirect(mapping: Dict[str, str]):
    """Within the block, ``load(a)`` returns module ``mapping[a]`` (same C ABI compiled into another library).  Used by
    POD: ``pod_sm100`` contains the prefill and decode launchers plus the fused kernel."""
    prev = getattr(_redirect, "map", None)
    _redirect.map = dict(mapping)
    try:
        yield
    finally:
        _redirect.map = prev


def load(name: str) -> NativeModule:
    """Load (JIT-building if needed) a native module. Fails loudly."""
    rmap = getattr(_redirect, "map", None)
    if rmap:
        name = rmap.get(name, name)
    mod = _loaded.get(name)
    if mod is not None:
        return mod
    with _build_lock:
        mod = _loaded.get(name)
        if mod is not None:
            return mod
        spec = REGISTRY[name]
        if os.environ.get("FIB200_DISABLE_JIT", "0") == "1":
            if not spec.so_path.exists():
                raise MissingNativeModuleError(f"native module '{name}' not built and JIT disabled")
        elif have_nvcc():
            build_module(spec, verbose=os.environ.get("FIB200_JIT_VERBOSE", "0") == "1")
        elif not spec.so_path.exists():
            raise MissingNativeModuleError(
                f"native module '{name}' is not built ({spec.so_path}) and nvcc is not available"
            )
        mod = NativeModule(name, spec.so_path)
        _loaded[name] = mod
        return mod


def native_launch_count() -> int:
    """Total number of kernels launched so far by all loaded native modules (each .so counts the
    launches that go through its LaunchCfg helper)."""
    total = 0
    for mod in _loaded.values():
        try:
            f = mod._dll.fib200_launch_count
            f.restype = ctypes.c_longlong
            total += int(f())
        except AttributeError:
            pass
    return total


def current_stream_ptr(device: Optional[torch.device] = None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def module_accessor(name: str, doc: str = ""):
    """``get_<x>_module()`` of the reference's op files: returns the loaded native module ``name`` (arguments that select a
    JIT specialisation in the reference - dtypes, head dims, backends - are accepted and ignored: one library serves all)."""

    def get(*args, **kwargs):
        return load(name)

    get.__doc__ = doc or f"The native module '{name}' (reference: the JIT-module accessor of the same name)."
    return get


# ----------------------------------------------------------------------------
# Reference JIT API names (flashinfer/jit/__init__.py): submodules first (they import the machinery above), then the flat
# re-exports users write ``from flashinfer.jit import ...`` against.
# ----------------------------------------------------------------------------
from . import env  # noqa: E402,F401
from . import cubin_loader  # noqa: E402,F401
from . import cpp_ext  # noqa: E402,F401
from . import utils  # noqa: E402,F401
from . import core  # noqa: E402,F401
from .core import *  # noqa: E402,F401,F403
from .core import _USER_SPECS  # noqa: E402,F401
from .cubin_loader import setup_cubin_loader  # noqa: E402,F401
from . import attention  # noqa: E402,F401
from .attention import *  # noqa: E402,F401,F403
