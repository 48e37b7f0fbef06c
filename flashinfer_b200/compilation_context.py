"""Target-architecture bookkeeping (reference flashinfer/compilation_context.py:27-101).  One target: sm_100a."""
from __future__ import annotations

import os


class CompilationContext:
    TARGET = (10, "0a")

    def __init__(self) -> None:
        env = os.environ.get("FLASHINFER_CUDA_ARCH_LIST")
        if env and env.strip() not in ("10.0a", "10.0"):
            raise RuntimeError(f"flashinfer_b200 only targets sm_100a (FLASHINFER_CUDA_ARCH_LIST={env!r})")
        self.TARGET_CUDA_ARCHS = {self.TARGET}

    def get_nvcc_flags_list(self, supported_major_versions=None):
        if supported_major_versions is not None and 10 not in supported_major_versions:
            raise RuntimeError("no supported architecture: this build only targets compute capability 10.0a")
        return ["-gencode", "arch=compute_100a,code=sm_100a"]


current_compilation_context = CompilationContext()
