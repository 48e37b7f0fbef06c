"""GEMM family. Parity: reference flashinfer/gemm/gemm_base.py (mm_bf16 :485, bmm_bf16 :692, tgv_gemm_sm100 :1446, ...)."""
from .dense import mm_bf16, bmm_bf16, tgv_gemm_sm100, mm_fp16, linear  # noqa: F401
