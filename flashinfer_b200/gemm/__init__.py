"""GEMM family. Parity: reference flashinfer/gemm/gemm_base.py (mm_bf16 :485, bmm_bf16 :692, tgv_gemm_sm100 :1446, ...)."""
from .dense import mm_bf16, bmm_bf16, tgv_gemm_sm100, mm_fp16, linear, linear_gated_silu, interleave_gate_up  # noqa: F401
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


def __getattr__(name):
    # router GEMMs / tinygemm live in dsv3_ops (which imports this package): resolve lazily to avoid the import cycle
    if name in ("mm_M1_16_K7168_N256", "mm_M1_16_K7168_N128", "mm_M1_16_K6144_N256", "tinygemm_bf16"):
        from .. import dsv3_ops

        return getattr(dsv3_ops, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


from .grouped import group_gemm_mxfp4_nt_groupwise as group_gemm_mxfp8_mxfp4_nt_groupwise  # noqa: F401,E402  (reference name)


def is_cute_dsl_available() -> bool:
    return False  # the GEMMs here are hand-written CUDA; nothing depends on nvidia-cutlass-dsl

from .. import _alias  # noqa: E402

_alias.install(__name__, ['gemm_base', 'routergemm'])  # the reference's per-file module paths


from .. import jit as _jit_acc  # noqa: E402

get_gemm_module = _jit_acc.module_accessor("gemm_sm100")
get_gemm_sm100_module = _jit_acc.module_accessor("gemm_sm100")
get_cutlass_fp4_gemm_module = _jit_acc.module_accessor("gemm_blockscaled_sm100")
get_cutlass_mxfp8_gemm_module = _jit_acc.module_accessor("gemm_blockscaled_sm100")
get_deepgemm_sm100_module = _jit_acc.module_accessor("gemm_blockscaled_sm100")
