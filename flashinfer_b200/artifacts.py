"""Binary artifacts of the build (reference flashinfer/artifacts.py).

The reference's artifacts are NVIDIA-built cubins and headers (trtllm-gen FMHA / batched GEMM / MoE, DeepGEMM) downloaded from an
artifact server and verified against published SHA-256 sums.  This library has no such inputs: every kernel is compiled from ``csrc/``.
Its only binary artifacts are the OUTPUTS of that build - ``_lib/<module>.so`` with the content hash of its sources, headers and flags
next to it - so the artifact vocabulary is mapped onto them: the "checksum" of a module is that content hash, an artifact is "present"
when its library exists and its hash matches the current sources, and "downloading" means building what is missing or stale."""
from __future__ import annotations

import os
from contextlib import contextmanager
from dataclasses import dataclass
from typing import Dict, Generator, List, Tuple


@contextmanager
def temp_env_var(key: str, value: str):
    """Set an environment variable inside the block and restore the previous state afterwards."""
    missing = object()
    old = os.environ.get(key, missing)
    os.environ[key] = value
    try:
        yield
    finally:
        if old is missing:
            os.environ.pop(key, None)
        else:
            os.environ[key] = old


def _family(name: str) -> str:
    if name.startswith(("prefill", "decode_sm100", "mla", "pod", "attention")):
        return "attention"
    if name.startswith(("gemm", "grouped_gemm", "decode_linear")):
        return "gemm"
    if name.startswith("comm"):
        return "comm"
    if name in ("moe",):
        return "moe"
    return "ops"


@dataclass(frozen=True)
class ArtifactPath:
    """Sub-directories of the reference's artifact server; here: families of native modules (``ArtifactPath.ATTENTION`` ...).  The
    reference's cubin families have no counterpart (empty strings)."""
    TRTLLM_GEN_FMHA: str = ""
    TRTLLM_GEN_BMM: str = ""
    TRTLLM_GEN_GEMM: str = ""
    DEEPGEMM: str = ""
    ATTENTION: str = "attention"
    GEMM: str = "gemm"
    MOE: str = "moe"
    COMM: str = "comm"
    OPS: str = "ops"


class CheckSumHash:
    """Expected checksums: the content hash (sources + included headers + flags) every registered module must carry."""

    @staticmethod
    def expected() -> Dict[str, str]:
        from . import jit

        return {name: spec.content_hash() for name, spec in jit.REGISTRY.items()}

    @staticmethod
    def recorded() -> Dict[str, str]:
        from . import jit

        out = {}
        for name, spec in jit.REGISTRY.items():
            if spec.hash_path.exists():
                out[name] = spec.hash_path.read_text().strip()
        return out


def get_checksums(subdirs=None) -> Dict[str, str]:
    """``{"<family>/<module>.so": expected content hash}``, optionally restricted to some families."""
    want = set(subdirs) if subdirs else None
    return {f"{_family(n)}/{n}.so": h for n, h in CheckSumHash.expected().items() if want is None or _family(n) in want}


def get_subdir_file_list() -> Generator[Tuple[str, str], None, None]:
    """``(relative artifact path, expected checksum)`` of every artifact of this build."""
    yield from sorted(get_checksums().items())


def get_available_cubin_files(source: str = "", retries: int = 3, delay: int = 5, timeout: int = 10) -> List[str]:
    """Libraries present in the tree (fresh or stale), as ``<family>/<module>.so``."""
    from . import jit

    return sorted(f"{_family(n)}/{n}.so" for n, spec in jit.REGISTRY.items() if spec.so_path.exists())


def get_available_header_files(source: str = "", retries: int = 3, delay: int = 5, timeout: int = 10) -> List[str]:
    """Headers the kernels are built against: the package's own include tree (nothing is fetched)."""
    from . import jit

    out = []
    for inc in jit.INCLUDE_DIRS:
        out += [str(p.relative_to(inc)) for p in sorted(inc.rglob("*")) if p.is_file()]
    return out


def get_artifacts_status() -> Tuple[Tuple[str, bool], ...]:
    """``(artifact, up to date?)`` for every module: present AND built from the current sources."""
    from . import jit

    return tuple((f"{_family(n)}/{n}.so", spec.is_fresh()) for n, spec in sorted(jit.REGISTRY.items()))


def download_artifacts() -> bool:
    """Make every artifact present: build the modules that are missing or stale (needs nvcc).  Returns True when all are fresh."""
    from . import jit

    stale = [n for n, spec in jit.REGISTRY.items() if not spec.is_fresh()]
    if stale and jit.have_nvcc():
        for n in stale:
            jit.build_module(jit.REGISTRY[n])
    return all(ok for _, ok in get_artifacts_status())


def clear_cubin() -> None:
    """Remove the built libraries (and their hashes); the next ``build`` / first use recompiles them."""
    from . import jit

    for spec in jit.REGISTRY.values():
        for p in (spec.so_path, spec.hash_path):
            if p.exists():
                p.unlink()
