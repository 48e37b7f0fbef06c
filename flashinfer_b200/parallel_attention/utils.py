"""Layout conversion, shard construction and group helpers (reference flashinfer/parallel_attention/utils.py)."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import torch
import torch.distributed as dist

_LAYOUTS = ("HND", "NHD")


def _check_layout(layout: str) -> None:
    if layout not in _LAYOUTS:
        raise ValueError(f"Invalid tensor layout: {layout}")


def convert_qkv_layout(q, k, v, src_layout: str, dst_layout: str):
    """``[H, S, D]`` (HND) <-> ``[S, H, D]`` (NHD) for q, k, v (contiguous results)."""
    _check_layout(src_layout), _check_layout(dst_layout)
    if src_layout == dst_layout:
        return q, k, v
    return tuple(t.transpose(0, 1).contiguous() for t in (q, k, v))


def convert_output_layout(out, src_layout: str, dst_layout: str):
    _check_layout(src_layout), _check_layout(dst_layout)
    return out if src_layout == dst_layout else out.transpose(0, 1).contiguous()


def _comm_device() -> torch.device:
    """Where small metadata tensors of collectives live: the current CUDA device under NCCL, the host otherwise."""
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def _as_lens(x) -> torch.Tensor:
    return x.to("cpu", torch.int64) if isinstance(x, torch.Tensor) else torch.tensor(list(x), dtype=torch.int64)


def split_varlen_input(tensor: torch.Tensor, seq_len_list, world_size: int, rank: int, tensor_layout: str = "HND") -> torch.Tensor:
    """Shard of ``rank`` for ring-parallel packed sequences: every sequence is cut into ``world_size`` chunks of
    ``ceil(len / world_size)`` tokens (the last rank takes what is left, possibly nothing), the rank's chunks of all sequences are
    concatenated, and the result is zero-padded to ``sum(ceil(len / world_size))`` so that all ranks hold the same shape."""
    _check_layout(tensor_layout)
    dim = 0 if tensor_layout == "NHD" else 1
    lens = _as_lens(seq_len_list)
    per = (lens + world_size - 1) // world_size
    pieces, offset = [], 0
    for n, base in zip(lens.tolist(), per.tolist()):
        lo = min(base * rank, n)
        hi = min(lo + base, n) if rank < world_size - 1 else n
        pieces.append(tensor.narrow(dim, offset + lo, max(hi - lo, 0)))
        offset += n
    shard = torch.cat(pieces, dim=dim)
    want = int(per.sum())
    if shard.shape[dim] < want:
        pad = list(shard.shape)
        pad[dim] = want - shard.shape[dim]
        shard = torch.cat([shard, shard.new_zeros(pad)], dim=dim)
    return shard


def _cumulative(lens: torch.Tensor, device) -> torch.Tensor:
    return torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)]).to(torch.int32).to(device)


def ulysses_varlen_config(seq_lens_q, seq_lens_kv) -> Tuple[torch.Tensor, torch.Tensor, int, int]:
    """``(cu_seqlens_q, cu_seqlens_kv, max_seqlen_q, max_seqlen_kv)`` of the whole packed batch (Ulysses-only varlen mode)."""
    lq, lk = _as_lens(seq_lens_q), _as_lens(seq_lens_kv)
    dev = _comm_device()
    return _cumulative(lq, dev), _cumulative(lk, dev), int(lq.max()), int(lk.max())


def ring_varlen_config(seq_lens_q, seq_lens_kv, ring_group) -> Tuple[torch.Tensor, torch.Tensor, int, int]:
    """Per-rank boundaries for ring-only varlen mode: ``(cu_seqlens_q [ring_size, n + 1], cu_seqlens_kv [ring_size, n + 1],
    max per-rank q chunk, max per-rank kv chunk)`` matching the shards :func:`split_varlen_input` builds."""
    world = dist.get_world_size(ring_group) if ring_group is not None else 1
    dev = _comm_device()

    def rows(lens: torch.Tensor):
        per = (lens + world - 1) // world
        out = []
        for r in range(world):
            mine = (lens - per * r).clamp(min=0)          # what is left of every sequence when rank r takes its turn
            if r < world - 1:
                mine = torch.minimum(mine, per)
            out.append(_cumulative(mine, dev))
        return torch.stack(out), int(per.max())

    cq, mq = rows(_as_lens(seq_lens_q))
    ck, mk = rows(_as_lens(seq_lens_kv))
    return cq, ck, mq, mk


def uneven_cp_config(seq_len: int, seq_len_padded: int, seq_len_cur_rank: int, ulysses_group=None, ring_group=None) -> Optional[torch.Tensor]:
    """Collective: every rank contributes its real token count; returns the real token count of every rank of the caller's ring group
    (a ring rank that is a Ulysses group counts the sum over that group), or ``None`` without ring parallelism.  Feed the result to
    :class:`UnevenCPConfig`."""
    dev = _comm_device()
    mine = torch.tensor([int(seq_len_cur_rank)], dtype=torch.int32, device=dev)
    everyone = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(everyone, mine)
    lens = torch.cat(everyone).cpu()
    ring_size = dist.get_world_size(ring_group) if ring_group is not None else 1
    if ring_size == 1:
        return None
    if ulysses_group is None or dist.get_world_size(ulysses_group) == 1:
        return lens[torch.tensor(dist.get_process_group_ranks(ring_group))]
    part = lens[torch.tensor(dist.get_process_group_ranks(ulysses_group))].sum().to(torch.int32).reshape(1).to(dev)
    ring = [torch.empty_like(part) for _ in range(ring_size)]
    dist.all_gather(ring, part, group=ring_group)
    return torch.cat(ring).cpu()


def get_parallel_groups(ulysses_size: int, ring_size: int, device_type: str = "cuda"):
    """``(ring_group, ulysses_group)`` of the calling rank (``None`` for a dimension of size 1).  Ranks are laid out as
    ``[replica][ring][ulysses]`` with Ulysses fastest: a Ulysses group is a run of consecutive ranks (its all-to-all stays inside one
    NVSwitch island), ring neighbours are ``ulysses_size`` apart.  Collective: every rank creates every group."""
    world = dist.get_world_size()
    span = ulysses_size * ring_size
    if world % span:
        raise ValueError(f"World size ({world}) is not divisible by total parallel size ({span})")
    me = dist.get_rank()
    ring_group = ulysses_group = None
    for rep in range(world // span):
        base = rep * span
        if ulysses_size > 1:
            for r in range(ring_size):
                ranks = [base + r * ulysses_size + u for u in range(ulysses_size)]
                g = dist.new_group(ranks)
                if me in ranks:
                    ulysses_group = g
        if ring_size > 1:
            for u in range(ulysses_size):
                ranks = [base + r * ulysses_size + u for r in range(ring_size)]
                g = dist.new_group(ranks)
                if me in ranks:
                    ring_group = g
    return ring_group, ulysses_group
