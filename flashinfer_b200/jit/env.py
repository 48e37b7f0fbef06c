"""Directory constants under the reference's names (flashinfer/jit/env.py).  One in-tree build directory serves as JIT cache,
AOT directory and "cubin" directory: nothing is downloaded and nothing lives under ``~/.cache``, so a snapshot of the source tree
carries its binaries with it."""
import os
import pathlib

_PKG = pathlib.Path(__file__).resolve().parent.parent

FLASHINFER_BASE_DIR = pathlib.Path(os.environ.get("FLASHINFER_WORKSPACE_BASE", str(_PKG)))
FLASHINFER_WORKSPACE_DIR = _PKG
FLASHINFER_CSRC_DIR = _PKG / "csrc"
FLASHINFER_GEN_SRC_DIR = _PKG / "_lib" / "gen"          # generated headers (attention variants)
FLASHINFER_INCLUDE_DIR = FLASHINFER_CSRC_DIR / "include"
FLASHINFER_JIT_DIR = _PKG / "_lib"
FLASHINFER_AOT_DIR = FLASHINFER_JIT_DIR
FLASHINFER_CUBIN_DIR = FLASHINFER_JIT_DIR
FLASHINFER_DATA = _PKG
FLASHINFER_TVM_BINDING_DIR = None                       # plain C ABI + ctypes: no TVM-FFI binding layer
CUTLASS_INCLUDE_DIRS: list = []                         # the kernels are self-contained inline PTX: no CUTLASS headers needed
SPDLOG_INCLUDE_DIR = None


def has_flashinfer_jit_cache() -> bool:
    """True when every registered module has an up-to-date library in the tree (the role of the reference's flashinfer-jit-cache wheel)."""
    from . import REGISTRY

    return all(spec.is_fresh() for spec in REGISTRY.values())


def has_flashinfer_cubin() -> bool:
    """The reference ships pre-compiled cubins in a separate wheel; every kernel here is built from ``csrc/``."""
    return False
