"""Module path of the reference's ``flashinfer.triton`` package.  There is no Triton in this framework: the entry points below run
the native sm_100a kernels (``sm_constraint_gemm``: SM-constrained persistent GEMM on tcgen05; ``activation`` / ``norm`` / ``cascade``: the
elementwise kernels of csrc/elementwise with the Triton helpers' optional fp8 scale arguments)."""
from . import activation, cascade, norm, sm_constraint_gemm  # noqa: F401
from .sm_constraint_gemm import gemm, gemm_descriptor_persistent, gemm_persistent  # noqa: F401
