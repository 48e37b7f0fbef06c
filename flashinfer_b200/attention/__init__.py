"""Holistic batch attention.  Parity: reference flashinfer/attention/_core.py."""
from ._core import BatchAttention, BatchAttentionWithAttentionSinkWrapper, apply_attention_sink  # noqa: F401
