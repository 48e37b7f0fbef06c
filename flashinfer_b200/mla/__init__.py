"""Multi-head Latent Attention.  Parity: reference flashinfer/mla/_core.py."""
from ._core import (  # noqa: F401
    BatchMLAPagedAttentionWrapper,
    mla_attention_ref,
    trtllm_batch_decode_with_kv_cache_mla,
    xqa_batch_decode_with_kv_cache_mla,
)
from ._core import (  # noqa: F401,E402
    MLAHeadDimensions,
    MLALayerDimensions,
    deepseek_mla_dimensions,
    smaller_mla_dimensions,
    supported_mla_layer_dimensions,
)


from .. import jit as _jit_acc  # noqa: E402

get_mla_module = _jit_acc.module_accessor("mla_sm100")
get_batch_mla_module = _jit_acc.module_accessor("mla_sm100")
get_trtllm_gen_fmha_module = _jit_acc.module_accessor("mla_sm100")
