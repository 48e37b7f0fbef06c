"""Holistic batch attention.  Parity: reference flashinfer/attention/_core.py."""
from ._core import BatchAttention, BatchAttentionWithAttentionSinkWrapper, apply_attention_sink  # noqa: F401


from .. import jit as _jit_acc  # noqa: E402

get_holistic_attention_module = _jit_acc.module_accessor("prefill_sm100")
