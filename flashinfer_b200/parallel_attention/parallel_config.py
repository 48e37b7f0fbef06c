"""Configuration objects of context-parallel attention (reference flashinfer/parallel_attention/parallel_config.py).

Conventions (the reference's): every rank holds a shard of EQUAL length - the sequence is padded to a multiple of the world size and
the padding sits at the end of the last rank's shard.  The configs tell the wrappers how many tokens are real, so that padded keys
are dropped before they are attended to and padded output rows come back as zeros."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist


def _size(group) -> int:
    return dist.get_world_size(group) if group is not None else 1


@dataclass
class UnevenCPConfig:
    """One sequence whose length is not a multiple of the world size.

    ``seq_len``: real total length; ``seq_len_padded``: padded total (divisible by the world size);
    ``seq_len_cur_ring_group``: real token count of every rank of this rank's ring group (``[ring_size]``; with Ulysses inside the ring
    the count of a ring rank is the sum over its Ulysses group) - what :func:`~flashinfer_b200.parallel_attention.utils.uneven_cp_config`
    returns; ``None`` without ring parallelism."""
    seq_len: Optional[int] = None
    seq_len_padded: Optional[int] = None
    seq_len_cur_ring_group: Optional[torch.Tensor] = None

    def reset(self) -> None:
        self.seq_len = self.seq_len_padded = self.seq_len_cur_ring_group = None


@dataclass
class VarlenCPConfig:
    """Several sequences packed along the token dimension.  Two mutually exclusive modes:

    * Ulysses only: the packed batch is exchanged as a whole, the kernel needs the global ``cu_seqlens`` (``*_cur_ulysses_group``);
    * Ring only: every sequence is cut into ``ring_size`` chunks, rank r holds chunk r of every sequence; ``cu_seqlens`` are 2-d
      ``[ring_size, num_seqs + 1]``, row r = the boundaries inside rank r's shard (``*_cur_ring_group``)."""
    cu_seqlens_q_cur_ulysses_group: Optional[torch.Tensor] = None
    cu_seqlens_kv_cur_ulysses_group: Optional[torch.Tensor] = None
    max_seq_len_q_cur_ulysses_group: Optional[int] = None
    max_seq_len_kv_cur_ulysses_group: Optional[int] = None
    cu_seqlens_q_cur_ring_group: Optional[torch.Tensor] = None
    cu_seqlens_kv_cur_ring_group: Optional[torch.Tensor] = None
    max_seq_len_q_cur_ring_group: Optional[int] = None
    max_seq_len_kv_cur_ring_group: Optional[int] = None

    def set_varlen_cp_config(self, cu_seqlens_q_all_ranks, cu_seqlens_kv_all_ranks, max_seq_len_q, max_seq_len_kv, ulysses_group, ring_group) -> None:
        """Store the outputs of ``ulysses_varlen_config`` / ``ring_varlen_config`` in the fields of the active mode."""
        if _size(ring_group) == 1:
            self.cu_seqlens_q_cur_ulysses_group, self.cu_seqlens_kv_cur_ulysses_group = cu_seqlens_q_all_ranks, cu_seqlens_kv_all_ranks
            self.max_seq_len_q_cur_ulysses_group, self.max_seq_len_kv_cur_ulysses_group = max_seq_len_q, max_seq_len_kv
        elif _size(ulysses_group) == 1:
            self.cu_seqlens_q_cur_ring_group, self.cu_seqlens_kv_cur_ring_group = cu_seqlens_q_all_ranks, cu_seqlens_kv_all_ranks
            self.max_seq_len_q_cur_ring_group, self.max_seq_len_kv_cur_ring_group = max_seq_len_q, max_seq_len_kv
        else:
            raise NotImplementedError("Varlen CP only supported when ulysses_size == 1 or ring_size == 1")

    def reset(self) -> None:
        for name in self.__dataclass_fields__:
            setattr(self, name, None)
