"""``ParallelAttention``: one attention backend under Ulysses and / or Ring context parallelism (reference
flashinfer/parallel_attention/parallel_attention.py)."""
from __future__ import annotations

from typing import Optional

from .attention_ops import AttentionOpManager
from .parallel_config import UnevenCPConfig, VarlenCPConfig
from .parallel_wrapper import ring_wrapper, ulysses_wrapper


class ParallelAttention:
    """``run(query, key, value, tensor_layout)`` takes this rank's sequence shard (``[H, S, D]`` for HND, ``[S, H, D]`` for NHD) and
    returns the attention output of the same shard.  Ulysses runs inside ``ulysses_group`` (all-to-all: heads are scattered, the
    sequence gathered), Ring across ``ring_group`` (KV shards rotate, partial results merge by log-sum-exp); both together give 2-d
    parallelism.  ``uneven_cp_config`` / ``varlen_cp_config`` describe padded and packed inputs (see parallel_config)."""

    def __init__(self, attn_type: str = "sm100", ulysses_group=None, ring_group=None, uneven_cp_config: Optional[UnevenCPConfig] = None,
                 varlen_cp_config: Optional[VarlenCPConfig] = None, fuse_qkv: bool = False) -> None:
        self.attn_type = attn_type
        self.attn_impl = AttentionOpManager.get_impl(attn_type)
        self.ulysses_group, self.ring_group = ulysses_group, ring_group
        self.uneven_cp_config, self.varlen_cp_config = uneven_cp_config, varlen_cp_config
        self.fuse_qkv = fuse_qkv

    @ulysses_wrapper
    @ring_wrapper
    def run(self, query, key, value, tensor_layout, attn_mask=None, is_causal=False, return_lse=False, cur_rank_cu_seqlens_q=None,
            cur_rank_cu_seqlens_k=None, cur_rank_max_seqlen_q=0, cur_rank_max_seqlen_k=0, **kwargs):
        """``cur_rank_*`` and ``return_lse`` are set by the wrappers (sequence boundaries come from the configs); extra keyword
        arguments (e.g. ``sm_scale``) go to the backend."""
        return self.attn_impl(query=query, key=key, value=value, tensor_layout=tensor_layout, attn_mask=attn_mask, is_causal=is_causal,
                              return_lse=return_lse, cur_rank_cu_seqlens_q=cur_rank_cu_seqlens_q, cur_rank_cu_seqlens_k=cur_rank_cu_seqlens_k,
                              cur_rank_max_seqlen_q=cur_rank_max_seqlen_q, cur_rank_max_seqlen_k=cur_rank_max_seqlen_k, **kwargs)
