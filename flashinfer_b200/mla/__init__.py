"""Multi-head Latent Attention.  Parity: reference flashinfer/mla/_core.py."""
from ._core import (  # noqa: F401
    BatchMLAPagedAttentionWrapper,
    mla_attention_ref,
    trtllm_batch_decode_with_kv_cache_mla,
    xqa_batch_decode_with_kv_cache_mla,
)
