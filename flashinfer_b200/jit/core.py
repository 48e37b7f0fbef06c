"""The reference's JIT vocabulary (flashinfer/jit/core.py: JitSpec, JitSpecStatus, JitSpecRegistry, gen_jit_spec, build_jit_specs,
clear_cache_dir, the per-arch flag lists) on top of this package's ModuleSpec / build_module / load.  ``gen_jit_spec`` also lets
users compile their own .cu files with this toolchain (sm_100a flags, include path with fib200/*.cuh, uniform C ABI loader)."""
from __future__ import annotations

from enum import Enum as _Enum
from pathlib import Path
from types import SimpleNamespace as _NS
from typing import Dict, Optional, Sequence

from . import ARCH_FLAGS, REGISTRY, MissingNativeModuleError, ModuleSpec, NativeModule, _loaded, build_module, load  # noqa: F401
from ..compilation_context import current_compilation_context  # noqa: F401

MissingJITCacheError = MissingNativeModuleError

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
    for f in extra_cflags or []:                       # host-compiler flags travel through nvcc
        flags += ["-Xcompiler", str(f)]
    if needs_device_linking:
        flags.append("-rdc=true")
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


__all__ = ["JitSpec", "JitSpecStatus", "JitSpecRegistry", "jit_spec_registry", "gen_jit_spec", "build_jit_specs", "clear_cache_dir",
           "MissingJITCacheError", "current_compilation_context", "sm90a_nvcc_flags", "sm100a_nvcc_flags", "sm100f_nvcc_flags",
           "sm103a_nvcc_flags", "sm110a_nvcc_flags", "sm120a_nvcc_flags", "sm120f_nvcc_flags", "sm121a_nvcc_flags"]
