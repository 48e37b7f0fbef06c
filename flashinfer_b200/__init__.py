"""flashinfer_b200 — a B200-native (sm_100a) LLM-inference kernel library with FlashInfer's API surface."""
__version__ = "0.1.0"

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
