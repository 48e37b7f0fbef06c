"""``flashinfer.cudnn`` entry points (reference flashinfer/cudnn/{decode,prefill}.py).  There is no cuDNN graph behind
them here: they are the cuDNN call signatures in front of the native sm_100a attention kernels."""
from ..decode import cudnn_batch_decode_with_kv_cache
from ..prefill import cudnn_batch_prefill_with_kv_cache
from .. import _alias

__all__ = ["cudnn_batch_decode_with_kv_cache", "cudnn_batch_prefill_with_kv_cache"]
_alias.install(__name__, ["decode", "prefill", "utils"])
