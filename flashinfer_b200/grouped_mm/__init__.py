"""Parity: reference flashinfer/grouped_mm (grouped_mm_bf16 / fp8 / mxfp8 / fp4)."""
from ..gemm.grouped import grouped_mm_bf16  # noqa: F401
from ..gemm.grouped import grouped_mm_fp4, grouped_mm_fp8  # noqa: F401
from ..gemm.grouped import group_gemm_mxfp4_nt_groupwise as grouped_mm_mxfp4  # noqa: F401  (linear-scale MXFP4 form, extension)
from ..gemm.grouped import grouped_mm_mxfp8  # noqa: F401

from .. import _alias  # noqa: E402

_alias.install(__name__, ['core'])  # the reference's per-file module paths
