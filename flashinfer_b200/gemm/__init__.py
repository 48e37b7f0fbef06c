"""GEMM family. Parity: reference flashinfer/gemm/gemm_base.py (mm_bf16 :485, bmm_bf16 :692, tgv_gemm_sm100 :1446, ...)."""
from .dense import mm_bf16, bmm_bf16, tgv_gemm_sm100, mm_fp16, linear  # noqa: F401
from .grouped import (  # noqa: F401
    SegmentGEMMWrapper,
    batch_deepgemm_fp8_nt_groupwise,
    group_deepgemm_fp8_nt_groupwise,
    group_gemm_fp8_nt_groupwise,
    group_gemm_mxfp4_nt_groupwise,
    group_gemm_nvfp4_nt_groupwise,
    grouped_gemm_nt_masked,
    grouped_gemm_tiles,
    grouped_mm_bf16,
    segment_gemm,
)
from .lowp import (  # noqa: F401
    bmm_fp8,
    bmm_mxfp8,
    fp8_blockscale_gemm_sm90,
    gemm_fp8_nt_blockscaled,
    gemm_fp8_nt_groupwise,
    mm_fp4,
    mm_fp8,
    mm_mxfp8,
    prepare_low_latency_gemm_weights,
    trtllm_low_latency_gemm,
)
