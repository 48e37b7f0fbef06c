"""Sequence / context parallel attention (Ulysses all-to-all + Ring P2P).

Parity: reference flashinfer/parallel_attention/ (ParallelAttention, ulysses_wrapper / ring_wrapper
parallel_wrapper.py:10-527, UnevenCPConfig / VarlenCPConfig, get_parallel_groups).
Beyond the reference: the ring path supports causal attention (contiguous chunks: later-rank KV chunks are skipped,
the diagonal chunk runs the causal kernel).
"""
from .core import (  # noqa: F401
    ParallelAttention,
    UnevenCPConfig,
    VarlenCPConfig,
    get_parallel_groups,
    ring_attention,
    ring_varlen_config,
    split_varlen_input,
    ulysses_attention,
    ulysses_varlen_config,
    uneven_cp_config,
)

from .. import _alias  # noqa: E402

_alias.install(__name__, ['parallel_attention', 'parallel_config', 'parallel_wrapper', 'attention_ops', 'utils'])  # the reference's per-file module paths
