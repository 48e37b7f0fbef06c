from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from ..cascade import merge_state
from ..comm._p2p import all_to_all_uneven
from ..prefill import single_prefill_with_kv_cache


# ------------------------------------------------------------------ configs
@dataclass
class UnevenCPConfig:
    """Per-rank sequence lengths when the sequence does not divide evenly over the CP group."""
    seq_lens: Optional[List[int]] = None

    def set_uneven_cp_config(self, seq_lens: Sequence[int]) -> None:
        self.seq_lens = list(seq_lens)

    def reset(self) -> None:
        self.seq_lens = None


@dataclass
class VarlenCPConfig:
    """Packed variable-length batches: ``cu_seqlens`` of the *global* batch; every sequence is split evenly over the ranks."""
    cu_seqlens_q: Optional[torch.Tensor] = None
    cu_seqlens_k: Optional[torch.Tensor] = None
    max_seqlen_q: int = 0
    max_seqlen_k: int = 0

    def set_varlen_cp_config(self, cu_seqlens_q, cu_seqlens_k, max_seqlen_q: int = 0, max_seqlen_k: int = 0) -> None:
        self.cu_seqlens_q, self.cu_seqlens_k = cu_seqlens_q, cu_seqlens_k
        self.max_seqlen_q, self.max_seqlen_k = max_seqlen_q, max_seqlen_k

    def reset(self) -> None:
        self.cu_seqlens_q = self.cu_seqlens_k = None
        self.max_seqlen_q = self.max_seqlen_k = 0


def uneven_cp_config(total_len: int, world: int) -> UnevenCPConfig:
    base, rem = divmod(total_len, world)
    return UnevenCPConfig([base + (1 if r < rem else 0) for r in range(world)])


def split_varlen_input(x: torch.Tensor, cu_seqlens: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Take this rank's contiguous slice of every packed sequence (dim 0 = tokens)."""
    cu = cu_seqlens.tolist()
    parts = []
    for a, b in zip(cu[:-1], cu[1:]):
        n = b - a
        base, rem = divmod(n, world)
        lo = a + rank * base + min(rank, rem)
        parts.append(x[lo: lo + base + (1 if rank < rem else 0)])
    return torch.cat(parts, 0)


def _local_cu(cu_seqlens: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    cu = cu_seqlens.tolist()
    out = [0]
    for a, b in zip(cu[:-1], cu[1:]):
        base, rem = divmod(b - a, world)
        out.append(out[-1] + base + (1 if rank < rem else 0))
    return torch.tensor(out, dtype=torch.int32)


def ulysses_varlen_config(cu_seqlens_q, cu_seqlens_k, max_q: int = 0, max_k: int = 0) -> VarlenCPConfig:
    return VarlenCPConfig(cu_seqlens_q, cu_seqlens_k, max_q, max_k)


ring_varlen_config = ulysses_varlen_config


def get_parallel_groups(ulysses_size: int, ring_size: int, world_size: Optional[int] = None, rank: Optional[int] = None):
    """Ulysses groups are contiguous (intra-NVSwitch-island all-to-all), ring groups strided.  Returns
    ``(ulysses_group, ring_group)`` of the calling rank."""
    world_size = world_size or dist.get_world_size()
    rank = dist.get_rank() if rank is None else rank
    if ulysses_size * ring_size != world_size:
        raise ValueError("ulysses_size * ring_size must equal world_size")
    ug = rg = None
    for i in range(ring_size):
        ranks = list(range(i * ulysses_size, (i + 1) * ulysses_size))
        g = dist.new_group(ranks)
        if rank in ranks:
            ug = g
    for j in range(ulysses_size):
        ranks = list(range(j, world_size, ulysses_size))
        g = dist.new_group(ranks)
        if rank in ranks:
            rg = g
    return ug, rg


# ------------------------------------------------------------------ local attention
def _local_attn(q, k, v, causal: bool, sm_scale: Optional[float]):
    """NHD in, returns (o [S,H,D], lse [S,H] base-2)."""
    return single_prefill_with_kv_cache(q, k, v, causal=causal, sm_scale=sm_scale, return_lse=True)


# ------------------------------------------------------------------ Ulysses
def _a2a_seq_to_head(x: torch.Tensor, group, lens: Optional[List[int]]) -> torch.Tensor:
    """[S_local, H, D] -> [S_total, H/P, D]"""
    P = dist.get_world_size(group)
    if P == 1:
        return x
    S, H, D = x.shape
    xs = x.reshape(S, P, H // P, D).permute(1, 0, 2, 3).contiguous()  # [P, S, H/P, D]
    if lens is None:
        out = torch.empty_like(xs)
        dist.all_to_all_single(out, xs, group=group)
        return out.reshape(P * S, H // P, D)
    outs = [torch.empty(n, H // P, D, dtype=x.dtype, device=x.device) for n in lens]
    all_to_all_uneven(outs, [xs[i].contiguous() for i in range(P)], group)
    return torch.cat(outs, 0)


def _a2a_head_to_seq(o: torch.Tensor, group, lens: Optional[List[int]], rank: int) -> torch.Tensor:
    """[S_total, H/P, D] -> [S_local, H, D]"""
    P = dist.get_world_size(group)
    if P == 1:
        return o
    _, Hp, D = o.shape
    if lens is None:
        S = o.shape[0] // P
        xs = o.reshape(P, S, Hp, D).contiguous()
        out = torch.empty_like(xs)
        dist.all_to_all_single(out, xs, group=group)
        return out.permute(1, 0, 2, 3).reshape(S, P * Hp, D)
    chunks = list(torch.split(o, lens, 0))
    S = lens[rank]
    outs = [torch.empty(S, Hp, D, dtype=o.dtype, device=o.device) for _ in range(P)]
    all_to_all_uneven(outs, [c.contiguous() for c in chunks], group)
    return torch.stack(outs, 1).reshape(S, P * Hp, D)


def ulysses_attention(q, k, v, group, causal: bool = False, sm_scale: Optional[float] = None,
                      seq_lens: Optional[List[int]] = None, fuse_qkv: bool = False,
                      inner: Optional[Callable] = None):
    """DeepSpeed-Ulysses: all-to-all scatters heads / gathers sequence, full-sequence attention on H/P heads, and the
    reverse all-to-all on the output.  ``fuse_qkv`` sends Q, K, V in one all-to-all when their head counts match."""
    P = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if q.shape[1] % P or k.shape[1] % P:
        raise ValueError("Ulysses needs num_qo_heads and num_kv_heads divisible by the group size")
    if fuse_qkv and q.shape == k.shape == v.shape:
        fused = _a2a_seq_to_head(torch.cat([q, k, v], -1), group, seq_lens)
        D = q.shape[-1]
        qf, kf, vf = fused[..., :D].contiguous(), fused[..., D:2 * D].contiguous(), fused[..., 2 * D:].contiguous()
    else:
        qf, kf, vf = (_a2a_seq_to_head(t, group, seq_lens) for t in (q, k, v))
    if inner is not None:
        o = inner(qf, kf, vf)
    else:
        o, _ = _local_attn(qf, kf, vf, causal, sm_scale)
    return _a2a_head_to_seq(o, group, seq_lens, rank)


# ------------------------------------------------------------------ Ring
def ring_attention(q, k, v, group, causal: bool = False, sm_scale: Optional[float] = None, return_lse: bool = False):
    """Ring attention over contiguous sequence chunks: KV blocks rotate rank -> rank+1 with double-buffered
    isend/irecv overlapping the per-step attention; partial states are merged with the (o, lse) algebra."""
    P = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if P == 1:
        o, lse = _local_attn(q, k, v, causal, sm_scale)
        return (o, lse) if return_lse else o
    ranks = dist.get_process_group_ranks(group)
    nxt, prv = ranks[(rank + 1) % P], ranks[(rank - 1) % P]
    # chunk sizes may differ per rank: exchange the lengths once
    lens_t = torch.tensor([k.shape[0]], dtype=torch.int64, device=k.device)
    all_lens = [torch.zeros_like(lens_t) for _ in range(P)]
    dist.all_gather(all_lens, lens_t, group=group)
    all_lens = [int(t) for t in all_lens]
    cur = torch.stack([k, v]).contiguous()  # [2, S, Hkv, D]
    o_acc = lse_acc = None
    for step in range(P):
        src = (rank - step) % P
        reqs = []
        nxt_buf = None
        if step < P - 1:
            nsrc = (rank - step - 1) % P
            nxt_buf = torch.empty(2, all_lens[nsrc], k.shape[1], k.shape[2], dtype=k.dtype, device=k.device)
            ops = [dist.P2POp(dist.isend, cur, nxt, group), dist.P2POp(dist.irecv, nxt_buf, prv, group)]
            if rank % 2:
                ops.reverse()
            reqs = dist.batch_isend_irecv(ops)
        if not (causal and src > rank):
            o, lse = _local_attn(q, cur[0], cur[1], causal and src == rank, sm_scale)
            if o_acc is None:
                o_acc, lse_acc = o, lse
            else:
                o_acc, lse_acc = merge_state(o_acc, lse_acc, o, lse)
        for r in reqs:
            r.wait()
        if nxt_buf is not None:
            cur = nxt_buf
    return (o_acc, lse_acc) if return_lse else o_acc


# ------------------------------------------------------------------ unified front end
class ParallelAttention:
    """2-D sequence parallelism: Ulysses inside ``ulysses_group`` and Ring across ``ring_group``.

    ``run(query, key, value, tensor_layout)`` takes this rank's sequence shard (``[S, H, D]`` for NHD or ``[H, S, D]``
    for HND) and returns the attention output for the same shard."""

    def __init__(self, attn_type: str = "sm100", ulysses_group=None, ring_group=None,
                 uneven_cp_config: Optional[UnevenCPConfig] = None, varlen_cp_config: Optional[VarlenCPConfig] = None,
                 fuse_qkv: bool = False) -> None:
        self.attn_type = attn_type
        self.ulysses_group, self.ring_group = ulysses_group, ring_group
        self.uneven_cp_config, self.varlen_cp_config = uneven_cp_config, varlen_cp_config
        self.fuse_qkv = fuse_qkv

    def run(self, query, key, value, tensor_layout: str = "NHD", attn_mask=None, is_causal: bool = False,
            return_lse: bool = False, sm_scale: Optional[float] = None, **kwargs):
        if attn_mask is not None:
            raise NotImplementedError("attn_mask is not supported by parallel attention")
        if return_lse:
            raise ValueError("return_lse is managed internally by the ring wrapper")
        if self.varlen_cp_config is not None and self.varlen_cp_config.cu_seqlens_q is not None:
            raise NotImplementedError("packed varlen batches: call run() per sequence (split_varlen_input gives the shards)")
        hnd = tensor_layout == "HND"
        if hnd:
            query, key, value = (t.transpose(0, 1).contiguous() for t in (query, key, value))
        lens = self.uneven_cp_config.seq_lens if self.uneven_cp_config and self.uneven_cp_config.seq_lens else None
        ring, uly = self.ring_group, self.ulysses_group
        ring_on = ring is not None and dist.get_world_size(ring) > 1
        if is_causal and ring_on and uly is not None and dist.get_world_size(uly) > 1:
            raise NotImplementedError("causal + combined Ulysses x Ring needs a zig-zag layout; use one of the two")
        inner = (lambda q, k, v: ring_attention(q, k, v, ring, is_causal, sm_scale)) if ring_on else None
        if uly is not None and dist.get_world_size(uly) > 1:
            out = ulysses_attention(query, key, value, uly, is_causal, sm_scale, lens, self.fuse_qkv, inner)
        elif inner is not None:
            out = inner(query, key, value)
        else:
            out, _ = _local_attn(query, key, value, is_causal, sm_scale)
        return out.transpose(0, 1).contiguous() if hnd else out
