"""Fused MoE.  Parity: reference flashinfer/fused_moe/ (core.py, fused_routing_dsv3.py, cute_dsl/)."""
from .core import (  # noqa: F401
    ActivationType,
    CuteDslMoEWrapper,
    GatedActType,
    RoutingMethodType,
    WeightLayout,
    b12x_fused_moe,
    B12xMoEWrapper,
    cute_dsl_fused_moe_nvfp4,
    cutlass_fused_moe,
    fused_topk_deepseek,
    moe_forward,
    moe_forward_nvfp4,
    moe_reference,
    reorder_rows_for_gated_act_gemm,
    route,
    trtllm_bf16_moe,
    trtllm_bf16_routed_moe,
    trtllm_fp4_block_scale_moe,
    trtllm_fp4_block_scale_routed_moe,
    trtllm_fp8_block_scale_moe,
    trtllm_fp8_block_scale_routed_moe,
    trtllm_fp8_per_tensor_scale_moe,
    trtllm_mxint4_block_scale_moe,
)
from .core import Fp8QuantizationType, MoEInputs, RoutingInputMode, moe_forward_fp8_block  # noqa: F401,E402


def convert_to_block_layout(input_tensor, blockK: int):
    """``[..., M, K]`` -> ``[..., K / blockK, M, blockK]`` (the BlockMajorK weight layout of the reference's trtllm-gen MoE)."""
    *lead, M, K = input_tensor.shape
    return input_tensor.reshape(*lead, M, K // blockK, blockK).transpose(-3, -2).contiguous()

from .. import _alias  # noqa: E402

_alias.install(__name__, ['fused_routing_dsv3'])  # the reference's per-file module paths
