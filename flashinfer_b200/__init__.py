"""flashinfer_b200 — a B200-native (sm_100a) LLM-inference kernel library with FlashInfer's API surface."""
from .version import __version__  # noqa: F401

from . import jit, reference, utils  # noqa: F401
from .decode import (  # noqa: F401
    BatchDecodeWithPagedKVCacheWrapper,
    CUDAGraphBatchDecodeWithPagedKVCacheWrapper,
    fast_decode_plan,
    single_decode_with_kv_cache,
)
from .gemm import bmm_bf16, mm_bf16, tgv_gemm_sm100  # noqa: F401
from .utils import MaskMode, PosEncodingMode, TensorLayout, next_positive_power_of_2  # noqa: F401
from . import activation, cascade, norm, page, rope  # noqa: F401,E402
from .activation import gelu_and_mul, gelu_tanh_and_mul, silu_and_mul  # noqa: F401,E402
from .cascade import merge_state, merge_state_in_place, merge_states  # noqa: F401,E402
from .norm import (  # noqa: F401,E402
    fused_add_rmsnorm,
    fused_add_rmsnorm_quant,
    fused_rmsnorm_silu,
    gemma_fused_add_rmsnorm,
    gemma_rmsnorm,
    layernorm,
    rmsnorm,
    rmsnorm_quant,
)
from .page import (  # noqa: F401,E402
    append_paged_kv_cache,
    append_paged_mla_kv_cache,
    get_batch_indices_positions,
    get_seq_lens,
)
from .prefill import (  # noqa: F401,E402
    BatchPrefillWithPagedKVCacheWrapper,
    BatchPrefillWithRaggedKVCacheWrapper,
    single_prefill_with_kv_cache,
    single_prefill_with_kv_cache_return_lse,
)
from .rope import (  # noqa: F401,E402
    apply_llama31_rope,
    apply_llama31_rope_inplace,
    apply_llama31_rope_pos_ids,
    apply_llama31_rope_pos_ids_inplace,
    apply_rope,
    apply_rope_inplace,
    apply_rope_pos_ids,
    apply_rope_pos_ids_inplace,
    apply_rope_with_cos_sin_cache,
    apply_rope_with_cos_sin_cache_inplace,
)
from . import quantization, sampling, topk  # noqa: F401,E402
from .quantization import (  # noqa: F401,E402
    SfLayout,
    block_scale_interleave,
    e2m1_and_ufp8sf_scale_to_float,
    fp4_quantize,
    mxfp4_dequantize,
    mxfp4_dequantize_host,
    mxfp4_quantize,
    mxfp8_dequantize_host,
    mxfp8_quantize,
    nvfp4_batched_quantize,
    nvfp4_block_scale_interleave,
    nvfp4_kv_dequantize,
    nvfp4_kv_quantize,
    nvfp4_quantize,
    packbits,
    scaled_fp4_grouped_quantize,
    segment_packbits,
    shuffle_matrix_a,
    shuffle_matrix_sf_a,
)
from .sampling import (  # noqa: F401,E402
    chain_speculative_sampling,
    min_p_sampling_from_probs,
    sampling_from_logits,
    sampling_from_probs,
    softmax,
    top_k_mask_logits,
    top_k_renorm_probs,
    top_k_sampling_from_probs,
    top_k_top_p_sampling_from_logits,
    top_k_top_p_sampling_from_probs,
    top_p_renorm_probs,
    top_p_sampling_from_probs,
)
from .topk import TopKTieBreak, top_k, top_k_page_table_transform, top_k_ragged_transform  # noqa: F401,E402
from . import fused_moe, grouped_mm  # noqa: F401,E402
from .fused_moe import (  # noqa: F401,E402
    ActivationType,
    GatedActType,
    RoutingMethodType,
    WeightLayout,
    cutlass_fused_moe,
    fused_topk_deepseek,
    reorder_rows_for_gated_act_gemm,
    trtllm_bf16_moe,
    trtllm_bf16_routed_moe,
    trtllm_fp4_block_scale_moe,
    trtllm_fp4_block_scale_routed_moe,
    trtllm_fp8_block_scale_moe,
    trtllm_fp8_block_scale_routed_moe,
    trtllm_fp8_per_tensor_scale_moe,
    trtllm_mxint4_block_scale_moe,
)
from .gemm import SegmentGEMMWrapper, grouped_mm_bf16  # noqa: F401,E402
from .gemm import bmm_fp8, bmm_mxfp8, gemm_fp8_nt_groupwise, mm_fp4, mm_fp8, mm_mxfp8  # noqa: F401,E402
from . import api_logging, autotuner, fi_trace, green_ctx, logits_processor, parallel_attention, profiler, testing, trace  # noqa: F401,E402
from .autotuner import autotune  # noqa: F401,E402
from .fi_trace import fi_trace  # noqa: F401,E402,F811  (the function shadows the module attribute, as in the reference)
from .api_logging import flashinfer_api  # noqa: F401,E402
from . import comm, mla, attention  # noqa: F401,E402
from . import concat_ops, diffusion_ops, dsv3_ops, gdn, mamba  # noqa: F401,E402
from .gdn import chunk_gated_delta_rule  # noqa: F401,E402
from .activation import silu_and_mul_scaled_nvfp4_experts_quantize  # noqa: F401,E402
from .norm import (  # noqa: F401,E402
    add_rmsnorm_fp4quant,
    fused_dit_gate_residual_layernorm_gamma_beta,
    fused_dit_gate_residual_layernorm_scale_shift,
    fused_dit_residual_layernorm_scale_shift,
    rmsnorm_fp4quant,
)
from .quantization.fp4 import nvfp4_quantize_paged_kv_cache  # noqa: F401,E402
from .gemm import prepare_low_latency_gemm_weights, trtllm_low_latency_gemm  # noqa: F401,E402
# ---- the rest of the reference's top-level namespace (flashinfer/__init__.py) ----
from .attention import BatchAttention, BatchAttentionWithAttentionSinkWrapper  # noqa: F401,E402
from .cascade import (  # noqa: F401,E402
    BatchDecodeWithSharedPrefixPagedKVCacheWrapper,
    BatchPrefillWithSharedPrefixPagedKVCacheWrapper,
    MultiLevelCascadeAttentionWrapper,
)
from .decode import BatchDecodeMlaWithPagedKVCacheWrapper, cudnn_batch_decode_with_kv_cache  # noqa: F401,E402
from .fused_moe.core import B12xMoEWrapper, CuteDslMoEWrapper, b12x_fused_moe, cute_dsl_fused_moe_nvfp4  # noqa: F401,E402
from .grouped_mm import grouped_mm_fp4, grouped_mm_fp8, grouped_mm_mxfp8  # noqa: F401,E402
from .mla import BatchMLAPagedAttentionWrapper  # noqa: F401,E402
from .pod import BatchPODWithPagedKVCacheWrapper, PODWithPagedKVCacheWrapper  # noqa: F401,E402
from .prefill import trtllm_fmha_v2_prefill  # noqa: F401,E402
from .sparse import BlockSparseAttentionWrapper, VariableBlockSparseAttentionWrapper  # noqa: F401,E402
from .xqa import xqa, xqa_mla  # noqa: F401,E402


def get_fp4_quantization_module(backend: str = "100"):
    """Reference fp4_quantization.py: returns the JIT module object behind the FP4 quantisers; here the native library."""
    return jit.load("quantization")


api_logging.instrument()  # @flashinfer_api on every public op and wrapper plan / run (no-op wrappers at FLASHINFER_LOGLEVEL=0)
trace.attach()  # bind op templates (wraps the ops only when FLASHINFER_TRACE_DUMP / FLASHINFER_TRACE_DIR is set)
