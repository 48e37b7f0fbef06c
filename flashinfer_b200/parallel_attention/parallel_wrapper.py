"""Ulysses (all-to-all over heads) and Ring (peer-to-peer KV rotation with log-sum-exp merging) as decorators around an attention call
(reference flashinfer/parallel_attention/parallel_wrapper.py).

Data conventions: q / k / v are this rank's shard, ``[H, S, D]`` (HND) or ``[S, H, D]`` (NHD), equal shapes on every rank (padding at
the end, see parallel_config).  Differences from the reference: partial results are accumulated in fp32 and steps whose KV shard holds
no real token are skipped; causal masking is available for plain (no uneven / varlen config) Ulysses-only and Ring-only runs - shards
are contiguous chunks in rank order, so a ring step with ``kv_rank > rank`` is skipped and the diagonal step runs the causal kernel."""
from __future__ import annotations

import functools
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def _dims(tensor_layout: str) -> Tuple[int, int]:
    """(sequence dim, head dim) of a 3-d q / k / v tensor."""
    if tensor_layout == "HND":
        return 1, 0
    if tensor_layout == "NHD":
        return 0, 1
    raise ValueError(f"Invalid tensor layout: {tensor_layout}")


def all_to_all(tensor: torch.Tensor, scatter_idx: int, gather_idx: int, tensor_layout: Optional[str] = None, group=None) -> torch.Tensor:
    """Split ``tensor`` into ``world_size`` pieces along ``scatter_idx``, send piece j to rank j, concatenate what arrives (in rank
    order) along ``gather_idx``."""
    if not dist.is_initialized():
        return tensor
    world = dist.get_world_size(group)
    if world == 1:
        return tensor
    if scatter_idx == gather_idx:
        raise ValueError("scatter_idx and gather_idx must be different")
    if tensor.shape[scatter_idx] % world:
        raise ValueError(f"Dimension {scatter_idx} of tensor {tuple(tensor.shape)} must be divisible by world size {world}")
    send = torch.stack(tensor.chunk(world, dim=scatter_idx)).contiguous()
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    return torch.cat(recv.unbind(0), dim=gather_idx).contiguous()


def ulysses_a2a_in(query, key, value, attn_mask, tensor_layout, ulysses_size=1, ulysses_rank=0, ulysses_group=None, fuse_qkv=False):
    """Shards ``[S / P, H, D]`` -> full sequence with ``H / P`` heads.  ``fuse_qkv`` sends q, k, v in one collective when their shapes
    agree (GQA shapes fall back to three collectives)."""
    if ulysses_size == 1:
        return query, key, value, attn_mask
    if attn_mask is not None:
        raise NotImplementedError("Attn mask not supported for ulysses_a2a_in")
    seq, head = _dims(tensor_layout)
    if fuse_qkv and query.shape == key.shape == value.shape:
        qkv = all_to_all(torch.stack([query, key, value]), head + 1, seq + 1, tensor_layout, ulysses_group)
        return qkv[0], qkv[1], qkv[2], attn_mask
    query, key, value = (all_to_all(t, head, seq, tensor_layout, ulysses_group) for t in (query, key, value))
    return query, key, value, attn_mask


def ulysses_a2a_out(output, tensor_layout, ulysses_size=1, ulysses_group=None):
    """Inverse exchange on the attention output: full sequence / ``H / P`` heads -> sequence shard / all heads."""
    if ulysses_size == 1:
        return output
    seq, head = _dims(tensor_layout)
    return all_to_all(output, seq, head, tensor_layout, ulysses_group)


def ring_fwd_out_correction(out: torch.Tensor, out_per_step: torch.Tensor, softmax_lse: torch.Tensor, softmax_lse_per_step: torch.Tensor) -> None:
    """In place: ``out <- (w out + w' out_per_step) / (w + w')`` with ``w = exp(softmax_lse)``, ``w' = exp(softmax_lse_per_step)``
    (natural log), written as a convex combination so that nothing overflows.  Call BEFORE :func:`ring_fwd_softmax_lse_correction`."""
    share = torch.sigmoid(softmax_lse_per_step - softmax_lse).nan_to_num(0.0).unsqueeze(-1)      # weight of the new block (nan: both empty)
    out.copy_(out + share.to(out.dtype) * (out_per_step.to(out.dtype) - out))


def ring_fwd_softmax_lse_correction(softmax_lse: torch.Tensor, softmax_lse_per_step: torch.Tensor) -> None:
    """In place: ``softmax_lse <- log(exp(softmax_lse) + exp(softmax_lse_per_step))``."""
    softmax_lse.copy_(torch.logaddexp(softmax_lse, softmax_lse_per_step.to(softmax_lse.dtype)))


def ring_attn_p2p_communicate(rank, send_tensor, send_dst, recv_tensor, recv_src, ring_group):
    """Post one send to ring rank ``send_dst`` and one receive from ``recv_src`` (group-local ranks); even ranks post the send first,
    odd ranks the receive, so the ring cannot deadlock on rendezvous transports.  Returns the requests."""
    send = dist.P2POp(dist.isend, send_tensor, group=ring_group, group_peer=send_dst)
    recv = dist.P2POp(dist.irecv, recv_tensor, group=ring_group, group_peer=recv_src)
    return dist.batch_isend_irecv([send, recv] if rank % 2 == 0 else [recv, send])


def get_kv_rank(ring_size: int, ring_rank: int, cur_iter: int) -> int:
    """Whose KV shard a rank holds at ring step ``cur_iter`` (shards travel rank -> rank + 1)."""
    return (ring_size + ring_rank - cur_iter) % ring_size


def _zero_tail(t: torch.Tensor, dim: int, start) -> None:
    start = int(start)
    if start < t.shape[dim]:
        t.narrow(dim, start, t.shape[dim] - start).zero_()


def ulysses_wrapper(func):
    @functools.wraps(func)
    def wrapper(self, query, key, value, tensor_layout, attn_mask=None, **kwargs):
        group, ring_group = self.ulysses_group, self.ring_group
        size = dist.get_world_size(group) if group is not None else 1
        ring_size = dist.get_world_size(ring_group) if ring_group is not None else 1
        if kwargs.get("return_lse", False):
            raise ValueError("return_lse=True is not supported in parallel attention")
        if size == 1:
            return func(self, query, key, value, tensor_layout, attn_mask, **kwargs)
        seq, head = _dims(tensor_layout)
        for name, t in (("query", query), ("key", key), ("value", value)):
            if t.shape[head] % size:
                raise ValueError(f"Head dim {head} of {name} {tuple(t.shape)} must be divisible by ulysses size {size}")
        uneven, varlen = self.uneven_cp_config, self.varlen_cp_config
        if kwargs.get("is_causal") and (uneven is not None or varlen is not None or ring_size > 1):
            raise NotImplementedError("causal parallel attention: plain Ulysses-only or Ring-only runs")
        query, key, value, attn_mask = ulysses_a2a_in(query, key, value, attn_mask, tensor_layout, size, dist.get_rank(group), group, self.fuse_qkv)
        real_q = None
        if ring_size == 1 and uneven is not None:                # the gathered sequence ends with the padding: drop padded keys
            real_q = int(uneven.seq_len)
            key, value = (t.narrow(seq, 0, real_q).contiguous() for t in (key, value))
        if ring_size == 1 and varlen is not None:
            cq, ck = varlen.cu_seqlens_q_cur_ulysses_group, varlen.cu_seqlens_kv_cur_ulysses_group
            kwargs.update(cur_rank_cu_seqlens_q=cq, cur_rank_cu_seqlens_k=ck, cur_rank_max_seqlen_q=varlen.max_seq_len_q_cur_ulysses_group,
                          cur_rank_max_seqlen_k=varlen.max_seq_len_kv_cur_ulysses_group)
            real_q = int(cq[-1])
            if key.shape[seq] != int(ck[-1]):
                key, value = (t.narrow(seq, 0, int(ck[-1])).contiguous() for t in (key, value))
        result = func(self, query, key, value, tensor_layout, attn_mask, **kwargs)
        if real_q is not None:
            _zero_tail(result, seq, real_q)                      # padded query rows come back as zeros
        return ulysses_a2a_out(result, tensor_layout, size, group)

    return wrapper


def ring_wrapper(func):
    @functools.wraps(func)
    def wrapper(self, query, key, value, tensor_layout, attn_mask=None, **kwargs):
        group = self.ring_group
        size = dist.get_world_size(group) if group is not None else 1
        if size == 1:
            return func(self, query, key, value, tensor_layout, attn_mask, **kwargs)
        rank = dist.get_rank(group)
        seq, _ = _dims(tensor_layout)
        uneven, varlen = self.uneven_cp_config, self.varlen_cp_config
        causal = bool(kwargs.get("is_causal"))
        if causal and (uneven is not None or varlen is not None):
            raise NotImplementedError("causal parallel attention: plain Ulysses-only or Ring-only runs")
        buffers: List[Optional[torch.Tensor]] = [torch.stack([key, value]).contiguous(), None]
        pending: List[list] = [[], []]
        out = lse = None
        for step in range(size):
            kv_rank = get_kv_rank(size, rank, step)
            for req in pending[(step + 1) % 2]:                  # the shard for this step was posted one step ago
                req.wait()
            pending[(step + 1) % 2] = []                         # (a second wait() on a finished gloo request never returns)
            if step < size - 1:                                  # pass the current shard on while it is being attended to
                buffers[(step + 1) % 2] = torch.empty_like(buffers[step % 2])
                pending[step % 2] = ring_attn_p2p_communicate(rank, buffers[step % 2], (rank + 1) % size, buffers[(step + 1) % 2],
                                                              (rank - 1) % size, group)
            kv = buffers[step % 2]
            step_kwargs = dict(kwargs, return_lse=True)
            if uneven is not None:
                kv = kv.narrow(seq + 1, 0, int(uneven.seq_len_cur_ring_group[kv_rank]))
            if varlen is not None:
                cq, ck = varlen.cu_seqlens_q_cur_ring_group[rank], varlen.cu_seqlens_kv_cur_ring_group[kv_rank]
                step_kwargs.update(cur_rank_cu_seqlens_q=cq, cur_rank_cu_seqlens_k=ck, cur_rank_max_seqlen_q=varlen.max_seq_len_q_cur_ring_group,
                                   cur_rank_max_seqlen_k=varlen.max_seq_len_kv_cur_ring_group)
                kv = kv.narrow(seq + 1, 0, int(ck[-1]))
            if causal:
                if kv_rank > rank:                               # a later chunk of the sequence: fully masked
                    continue
                step_kwargs["is_causal"] = kv_rank == rank
            if kv.shape[seq + 1] == 0:                           # that rank holds padding only
                continue
            o, l = func(self, query, kv[0].contiguous(), kv[1].contiguous(), tensor_layout, attn_mask, **step_kwargs)
            if out is None:
                out, lse = o.float(), l.float().clone()
            else:
                ring_fwd_out_correction(out, o, lse, l.float())
                ring_fwd_softmax_lse_correction(lse, l.float())
        if out is None:
            out = torch.zeros_like(query, dtype=torch.float32)
        out = out.to(query.dtype)
        if uneven is not None:
            _zero_tail(out, seq, uneven.seq_len_cur_ring_group[rank])
        if varlen is not None:
            _zero_tail(out, seq, varlen.cu_seqlens_q_cur_ring_group[rank][-1])
        return out

    return wrapper
