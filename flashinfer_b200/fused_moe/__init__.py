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

from . import moe_utils  # noqa: E402,F401  (standalone building blocks: moe_sort / moe_permute / moe_unpermute / moe_activation ...)
from .moe_utils import (  # noqa: E402,F401
    MoeActivationType,
    allocate_moe_sort_buffers,
    get_max_num_permuted_tokens,
    get_max_num_tiles,
    moe_activation,
    moe_geglu,
    moe_gelu,
    moe_output_memset,
    moe_output_memset_inplace,
    moe_permute,
    moe_relu,
    moe_silu,
    moe_sort,
    moe_swiglu,
    moe_unpermute,
)
from .. import _alias  # noqa: E402

_alias.install(__name__, ['fused_routing_dsv3'])  # the reference's per-file module paths


from .. import jit as _jit_acc  # noqa: E402

get_cutlass_fused_moe_module = _jit_acc.module_accessor("moe")
get_trtllm_moe_sm100_module = _jit_acc.module_accessor("moe")


def get_reorder_rows_for_gated_act_gemm_row_indices(x):
    """Row permutation that interleaves the two halves of a gated FC1 weight: [r0, r(M/2), r1, r(M/2+1), ...]
    (reference fused_moe/core.py:191; the native GEMM's SwiGLU epilogue reads gate / up from adjacent rows)."""
    import torch

    m = x.shape[0]
    if x.dim() != 2 or m % 2:
        raise ValueError("expected a 2-D weight with an even number of rows")
    return torch.stack([torch.arange(m // 2), torch.arange(m // 2, m)], dim=1).reshape(-1)


_w2_perm_cache = {}


def get_w2_permute_indices_with_cache(cache, dst_w2_weight, epilogue_tile_m: int, num_elts_per_sf=None):
    """Row permutation the reference applies to FC2 weights (or their scale factors) for its trtllm-gen epilogue tiles, memoised
    per (shape, tile).  The kernels here read FC2 weights row-major, so the shuffle is the one :func:`shuffle_matrix_a` defines
    and is only needed when weights are exchanged with that layout."""
    import torch

    from ..quantization.fp4 import _shuffle_row_indices

    cache = cache if cache is not None else _w2_perm_cache
    key = ("w2", tuple(dst_w2_weight.shape), epilogue_tile_m, num_elts_per_sf)
    if key not in cache:
        cache[key] = _shuffle_row_indices(dst_w2_weight.shape[0], epilogue_tile_m, dst_w2_weight.device).to(torch.long)
    return cache[key]


def is_trtllm_moe_supported(dtype_weights, dtype_act, quant_method=None) -> bool:
    """Which (weight, activation) dtype pairs the routed-MoE pipeline accepts on this device (reference fused_moe/core.py:100):
    bf16 x bf16, e4m3 x e4m3 (per-tensor or 128-block scales), e2m1 x e2m1 (NVFP4), mx-e2m1 x {mx-e2m1, mx-e4m3, bf16}."""
    import torch

    from ..tllm_enums import DtypeTrtllmGen as D

    if not torch.cuda.is_available() or torch.cuda.get_device_capability()[0] < 10:
        return False
    ok = {D.Bfloat16: {D.Bfloat16}, D.E4m3: {D.E4m3}, D.E2m1: {D.E2m1}, D.MxE2m1: {D.MxE2m1, D.MxE4m3, D.Bfloat16}}
    return dtype_act in ok.get(dtype_weights, set())
