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

_PKG = Path(__file__).resolve().parent
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


# ============================================================================
# Reference JIT API names (flashinfer/jit/__init__.py, core.py): the same
# concepts on top of ModuleSpec.  ``gen_jit_spec`` also lets users compile their
# own .cu files with this toolchain (arch flags, include path with fib200/*.cuh,
# uniform C ABI loader).
# ============================================================================
from enum import Enum as _Enum
from types import SimpleNamespace as _NS

MissingJITCacheError = MissingNativeModuleError

env = _NS(
    FLASHINFER_WORKSPACE_DIR=_PKG,
    FLASHINFER_JIT_DIR=LIB_DIR,
    FLASHINFER_GEN_SRC_DIR=CSRC,
    FLASHINFER_CSRC_DIR=CSRC,
    FLASHINFER_INCLUDE_DIR=CSRC / "include",
    FLASHINFER_AOT_DIR=LIB_DIR,
    FLASHINFER_CUBIN_DIR=LIB_DIR,
    CUTLASS_INCLUDE_DIRS=[],
    SPDLOG_INCLUDE_DIR=None,
)

sm100a_nvcc_flags = list(ARCH_FLAGS)
sm100f_nvcc_flags = ["-gencode", "arch=compute_100f,code=sm_100f"]
sm103a_nvcc_flags = ["-gencode", "arch=compute_103a,code=sm_103a"]
sm110a_nvcc_flags = ["-gencode", "arch=compute_110a,code=sm_110a"]
sm120a_nvcc_flags = ["-gencode", "arch=compute_120a,code=sm_120a"]
sm120f_nvcc_flags = ["-gencode", "arch=compute_120f,code=sm_120f"]
sm121a_nvcc_flags = ["-gencode", "arch=compute_121a,code=sm_121a"]
sm90a_nvcc_flags = ["-gencode", "arch=compute_90a,code=sm_90a"]


class JitSpecStatus(_Enum):
    NOT_COMPILED = 0
    COMPILED = 1
    STALE = 2


class JitSpec(ModuleSpec):
    """ModuleSpec under the reference's name, with its convenience methods."""

    @property
    def jit_library_path(self) -> Path:
        return self.so_path

    @property
    def aot_path(self) -> Path:
        return self.so_path

    @property
    def is_compiled(self) -> bool:
        return self.so_path.exists()

    @property
    def status(self) -> JitSpecStatus:
        if not self.so_path.exists():
            return JitSpecStatus.NOT_COMPILED
        return JitSpecStatus.COMPILED if self.is_fresh() else JitSpecStatus.STALE

    def build(self, verbose: bool = False, need_lock: bool = True) -> None:
        build_module(self, verbose=verbose)

    def load(self, so_path=None) -> "NativeModule":
        return load(self.name)

    def build_and_load(self) -> "NativeModule":
        return load(self.name)


_USER_SPECS: Dict[str, JitSpec] = {}


def gen_jit_spec(name: str, sources: Sequence, extra_cflags: Optional[Sequence[str]] = None,
                 extra_cuda_cflags: Optional[Sequence[str]] = None, extra_ldflags: Optional[Sequence[str]] = None,
                 extra_include_paths: Optional[Sequence] = None, needs_device_linking: bool = False) -> JitSpec:
    """Declare a native module from ``sources`` (paths relative to ``csrc/`` or absolute: user kernels are welcome).  The
    result builds with the sm_100a flags of this package and loads through the uniform C-ABI caller (``spec.build_and_load()``)."""
    flags = list(extra_cuda_cflags or [])
    for inc in extra_include_paths or []:
        flags += ["-I", str(inc)]
    spec = JitSpec(name, [str(s) for s in sources], extra_flags=flags, ldflags=list(extra_ldflags or []))
    REGISTRY[name] = spec
    _USER_SPECS[name] = spec
    return spec


class JitSpecRegistry:
    """View of the module registry (reference jit/core.py JitSpecRegistry)."""

    def register(self, spec: ModuleSpec) -> None:
        REGISTRY[spec.name] = spec

    def get_all_specs(self) -> Dict[str, ModuleSpec]:
        return dict(REGISTRY)

    def get_spec_status(self, name: str):
        spec = REGISTRY.get(name)
        if spec is None:
            return None
        return _NS(name=name, status=(JitSpecStatus.NOT_COMPILED if not spec.so_path.exists() else
                                      (JitSpecStatus.COMPILED if spec.is_fresh() else JitSpecStatus.STALE)),
                   library_path=spec.so_path, sources=spec.source_paths())

    def get_all_statuses(self):
        return [self.get_spec_status(n) for n in REGISTRY]

    def get_stats(self) -> Dict[str, int]:
        st = [s.status for s in self.get_all_statuses()]
        return {"total": len(st), "compiled": sum(x == JitSpecStatus.COMPILED for x in st),
                "not_compiled": sum(x == JitSpecStatus.NOT_COMPILED for x in st), "stale": sum(x == JitSpecStatus.STALE for x in st)}


jit_spec_registry = JitSpecRegistry()


def build_jit_specs(specs: Sequence[ModuleSpec], verbose: bool = False, skip_prebuilt: bool = True) -> None:
    for spec in specs:
        build_module(spec, verbose=verbose, force=not skip_prebuilt and False)


def clear_cache_dir() -> None:
    """Remove the libraries of user-declared modules (``gen_jit_spec``) and the in-process module cache; the package's own
    libraries are sources of truth for the in-tree build and are rebuilt by ``build_all(force=True)`` instead."""
    for name, spec in list(_USER_SPECS.items()):
        for p in (spec.so_path, spec.hash_path):
            if p.exists():
                p.unlink()
        _loaded.pop(name, None)


from .compilation_context import current_compilation_context  # noqa: E402,F401


cubin_loader = _NS(get_cubin=lambda *a, **k: b"", setup_cubin_loader=lambda *a, **k: None)


def setup_cubin_loader(*args, **kwargs) -> None:
    return None  # no pre-built cubins: every kernel is compiled from csrc/


_GEN_MAP = {
    "gen_act_and_mul_module": "activation", "gen_batch_attention_module": "prefill_sm100", "gen_batch_decode_mla_module": "mla_sm100",
    "gen_batch_decode_module": "decode_sm100", "gen_batch_mla_module": "mla_sm100", "gen_batch_pod_module": "pod_sm100",
    "gen_batch_prefill_module": "prefill_sm100", "gen_comm_alltoall_module": "comm_alltoall", "gen_cudnn_fmha_module": "prefill_sm100",
    "gen_customize_batch_decode_module": "attention_generic", "gen_customize_single_decode_module": "attention_generic",
    "gen_dcp_alltoall_module": "comm_collectives", "gen_dsv3_fused_routing_module": "moe", "gen_dsv3_router_gemm_module": "gemm_sm100",
    "gen_fmha_cutlass_sm100a_module": "prefill_sm100", "gen_fmha_v2_module": "prefill_sm100", "gen_fp4_kv_dequantization_module": "quantization",
    "gen_fp4_kv_quantization_module": "quantization", "gen_moe_alltoall_module": "comm_alltoall", "gen_moe_utils_module": "moe",
    "gen_pod_module": "pod_sm100", "gen_single_decode_module": "decode_sm100", "gen_single_prefill_module": "prefill_sm100",
    "gen_tinygemm2_module": "gemm_sm100", "gen_trtllm_comm_module": "comm_allreduce", "gen_trtllm_fmha_v2_sm120_module": "prefill_sm100",
    "gen_trtllm_gen_fmha_module": "decode_sm100", "gen_trtllm_mnnvl_comm_module": "comm_allreduce", "gen_vllm_comm_module": "comm_allreduce",
}


def _make_gen(fn_name: str, module: str):
    def gen(*args, **kwargs):
        """Reference JIT generator name: the template arguments select an instance there; here the native module covers them."""
        spec = REGISTRY[module]
        js = JitSpec(spec.name, list(spec.sources), list(spec.extra_flags), list(spec.ldflags), list(spec.deps))
        return js
    gen.__name__ = fn_name
    return gen


for _fn, _mod in _GEN_MAP.items():
    globals()[_fn] = _make_gen(_fn, _mod)


# ----------------------------------------------------------------------------
# User attention variants compiled INTO the tcgen05 prefill kernel.
# Reference: flashinfer/jit/attention/modules.py gen_customize_batch_prefill_module / gen_customize_single_prefill_module
# (:1189-1708) + include/flashinfer/attention/variant_helper.cuh (REGISTER_LOGITS_TRANSFORM / REGISTER_LOGITS_MASK).
# ----------------------------------------------------------------------------
_VARIANT_CTYPES = {"float": "float", "float32": "float", "double": "double", "half": "__half", "float16": "__half",
                   "bfloat16": "__nv_bfloat16", "nv_bfloat16": "__nv_bfloat16", "int32_t": "int32_t", "int32": "int32_t", "int": "int32_t",
                   "int64_t": "int64_t", "int64": "int64_t", "uint8_t": "uint8_t", "uint8": "uint8_t", "bool": "bool", "uint32_t": "uint32_t"}

_VARIANT_PRELUDE = """// generated by flashinfer_b200.jit.gen_customize_batch_prefill_module - do not edit
// A variant is a struct with two static device hooks evaluated for every logit in the softmax pass of the tcgen05 prefill kernel:
//   static __device__ float LogitsTransform(const VariantCtx& ctx, float logits, int kv_idx);   // logits = q.k * sm_scale (after soft-cap)
//   static __device__ bool  LogitsMask(const VariantCtx& ctx, int kv_idx);                       // false = masked out
// ctx carries batch_idx, qo_idx (row inside the request), qo_head_idx, kv_head_idx, qo_len, kv_len.  The additional tensors / scalars
// declared to the generator are visible under their names (tensors as typed const pointers, scalars as float).
// The reference's REGISTER_* macros are accepted for the two hooks this kernel implements.
#define REGISTER_LOGITS_TRANSFORM(params, logits, batch_idx, qo_idx, kv_idx, qo_head_idx, kv_head_idx, ...)                          \
  static __device__ __forceinline__ float LogitsTransform(const VariantCtx& ctx, float logits, int kv_idx) {                       \
    const int batch_idx = ctx.batch_idx, qo_idx = ctx.qo_idx, qo_head_idx = ctx.qo_head_idx, kv_head_idx = ctx.kv_head_idx;        \
    (void)batch_idx; (void)qo_idx; (void)qo_head_idx; (void)kv_head_idx;                                                           \
    __VA_ARGS__                                                                                                                    \
  }
#define REGISTER_LOGITS_MASK(params, batch_idx, qo_idx, kv_idx, qo_head_idx, kv_head_idx, ...)                                       \
  static __device__ __forceinline__ bool LogitsMask(const VariantCtx& ctx, int kv_idx) {                                           \
    const int batch_idx = ctx.batch_idx, qo_idx = ctx.qo_idx, qo_head_idx = ctx.qo_head_idx, kv_head_idx = ctx.kv_head_idx;        \
    (void)batch_idx; (void)qo_idx; (void)qo_head_idx; (void)kv_head_idx;                                                           \
    __VA_ARGS__                                                                                                                    \
  }
struct VariantDefaults {
  static __device__ __forceinline__ float LogitsTransform(const VariantCtx&, float logits, int) { return logits; }
  static __device__ __forceinline__ bool LogitsMask(const VariantCtx&, int) { return true; }
};
"""


class AttentionVariantSpec(JitSpec):
    """JitSpec of a prefill module with a compiled-in variant; remembers the declared extra arguments of ``run()``."""

    additional_tensor_names: Sequence[str] = ()
    additional_scalar_names: Sequence[str] = ()


def gen_customize_batch_prefill_module(backend, uri: str, dtype_q=None, dtype_kv=None, dtype_o=None, idtype=None, head_dim_qk: int = 128,
                                       head_dim_vo: int = 128, additional_tensor_names: Sequence[str] = (),
                                       additional_tensor_dtypes: Sequence[str] = (), additional_scalar_names: Sequence[str] = (),
                                       additional_scalar_dtypes: Sequence[str] = (), variant_name: str = "Variant",
                                       variant_decl: str = "", **kwargs) -> AttentionVariantSpec:
    """Compile ``variant_decl`` (C++: ``struct <variant_name> : VariantDefaults { ... }`` overriding ``LogitsTransform`` and / or
    ``LogitsMask``, see the generated header's prelude) into a private copy of the tcgen05 prefill kernel and return its spec;
    ``BatchPrefillWith{Paged,Ragged}KVCacheWrapper(..., jit_args=[uri, dtype_q, dtype_kv, dtype_o, idtype, head_dim_qk, head_dim_vo,
    tensor_names, tensor_dtypes, scalar_names, scalar_dtypes, variant_name, variant_decl])`` builds and uses it, the extra
    tensors / scalars follow ``q, kv`` in ``run()`` (at most 8 of each; scalars reach the device as float)."""
    if len(additional_tensor_names) > 8 or len(additional_scalar_names) > 8:
        raise ValueError("attention variants take at most 8 additional tensors and 8 additional scalars")
    if len(additional_tensor_names) != len(additional_tensor_dtypes) or len(additional_scalar_names) != len(additional_scalar_dtypes):
        raise ValueError("names / dtypes of the additional arguments must pair up")
    gen_dir = LIB_DIR / "gen"
    gen_dir.mkdir(parents=True, exist_ok=True)
    name = "prefill_variant_" + "".join(c if c.isalnum() or c == "_" else "_" for c in str(uri))
    lines = [_VARIANT_PRELUDE]
    for i, (n, dt) in enumerate(zip(additional_tensor_names, additional_tensor_dtypes)):
        ct = _VARIANT_CTYPES.get(str(dt).replace("torch.", ""), None)
        if ct is None:
            raise ValueError(f"unsupported additional tensor dtype {dt!r}")
        lines.append(f"#define {n} (reinterpret_cast<const {ct}*>(ctx.params->var_ptr[{i}]))")
    for i, n in enumerate(additional_scalar_names):
        lines.append(f"#define {n} (ctx.params->var_f[{i}])")
    lines.append(variant_decl)
    header = gen_dir / f"{name}.cuh"
    text = "\n".join(lines) + "\n"
    if not header.exists() or header.read_text() != text:
        header.write_text(text)
    flags = [f"-DFIB_PREFILL_VARIANT={variant_name}", f'-DFIB_PREFILL_VARIANT_HEADER="{header}"']
    spec = AttentionVariantSpec(name, ["attention/prefill_sm100.cu"], extra_flags=flags, deps=[str(header)])
    spec.additional_tensor_names = tuple(additional_tensor_names)
    spec.additional_scalar_names = tuple(additional_scalar_names)
    REGISTRY[name] = spec
    _USER_SPECS[name] = spec
    return spec


def gen_customize_single_prefill_module(backend, uri: str, *args, **kwargs) -> AttentionVariantSpec:
    """Single-request flavour: the same module (a single request is a batch of one)."""
    return gen_customize_batch_prefill_module(backend, uri, *args, **kwargs)


def _uri(prefix: str):
    def f(*args, **kwargs) -> str:
        parts = [str(a).replace("torch.", "") for a in args] + [f"{k}_{v}" for k, v in sorted(kwargs.items())]
        return "_".join([prefix] + parts)
    f.__name__ = f"get_{prefix}_uri"
    return f


get_single_decode_uri = _uri("single_decode")
get_single_prefill_uri = _uri("single_prefill")
get_batch_decode_uri = _uri("batch_decode")
get_batch_prefill_uri = _uri("batch_prefill")
get_batch_decode_mla_uri = _uri("batch_decode_mla")
get_batch_mla_uri = _uri("batch_mla")
get_batch_attention_uri = _uri("batch_attention")
get_pod_uri = _uri("pod")


def get_act_and_mul_cu_str(act_func_name: str, act_func_def: str) -> str:
    """Source of a custom gated activation in the style of csrc/elementwise/activation.cu (for ``gen_jit_spec``)."""
    return (f"// generated: {act_func_name}\\n#include <fib200/common.cuh>\\n__device__ __forceinline__ float {act_func_name}(float x)"
            f" {{ {act_func_def} }}\\n")


def module_accessor(name: str, doc: str = ""):
    """``get_<x>_module()`` of the reference's op files: returns the loaded native module ``name`` (arguments that select a
    JIT specialisation in the reference - dtypes, head dims, backends - are accepted and ignored: one library serves all)."""

    def get(*args, **kwargs):
        return load(name)

    get.__doc__ = doc or f"The native module '{name}' (reference: the JIT-module accessor of the same name)."
    return get
