"""Parallelism in one place: process-group construction, weight sharding helpers and the communication objects the engines use.

* tensor / expert parallel: :func:`init_tensor_parallel`, :func:`shard_rows` / :func:`shard_cols`, :func:`all_reduce_fp32` (what
  ``models.transformer`` uses); the flagship Llama engine replaces the all-reduce by the in-GEMM NVLink all-reduce
  (:class:`~flashinfer_b200.gemm.decode_linear.FusedLinearTP`) or the fused AR + residual + RMSNorm kernels of
  :class:`~flashinfer_b200.comm.TPCommunicator`;
* context parallel: :class:`~flashinfer_b200.parallel_attention.ParallelAttention` (Ulysses x Ring) and :func:`get_parallel_groups`;
* rank bookkeeping for TP x PP x EP x CP layouts: :class:`~flashinfer_b200.comm.Mapping`."""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from ..comm import Mapping, TPCommunicator  # noqa: F401
from ..parallel_attention import ParallelAttention, get_parallel_groups  # noqa: F401


def init_tensor_parallel(tp_size: Optional[int] = None):
    """The tensor-parallel group of the calling rank: consecutive blocks of ``tp_size`` ranks (default: the whole world).  Collective -
    every rank creates every group.  Returns ``None`` for ``tp_size == 1``."""
    world = dist.get_world_size()
    tp_size = tp_size or world
    if world % tp_size:
        raise ValueError(f"world size {world} is not divisible by tp_size {tp_size}")
    if tp_size == 1:
        return None
    mine = None
    for start in range(0, world, tp_size):
        ranks = list(range(start, start + tp_size))
        g = dist.new_group(ranks)
        if dist.get_rank() in ranks:
            mine = g
    return mine


def shard_rows(weight: torch.Tensor, rank: int, world: int, blocks: int = 1) -> torch.Tensor:
    """Column-parallel slice of a ``[out, in]`` weight: the rank's share of the output rows.  ``blocks`` > 1 for row-stacked weights
    (``[gate | up]``, ``[q | k | v]`` with equal blocks): every block is sharded on its own so that the slices stay aligned."""
    parts = weight.view(blocks, -1, weight.shape[-1])
    if parts.shape[1] % world:
        raise ValueError(f"{parts.shape[1]} rows per block are not divisible by {world} ranks")
    n = parts.shape[1] // world
    return parts[:, rank * n:(rank + 1) * n].reshape(-1, weight.shape[-1]).contiguous()


def shard_cols(weight: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Row-parallel slice of a ``[out, in]`` weight: the rank's share of the input columns (its output is a partial sum)."""
    if weight.shape[-1] % world:
        raise ValueError(f"{weight.shape[-1]} columns are not divisible by {world} ranks")
    n = weight.shape[-1] // world
    return weight[..., rank * n:(rank + 1) * n].contiguous()


def all_reduce_fp32(partial: torch.Tensor, group) -> torch.Tensor:
    """Sum of the ranks' partial results, accumulated in fp32 so that the value does not depend on the reduction order of the
    transport; returned in the input dtype.  ``group is None`` (no parallelism) returns the input."""
    if group is None or dist.get_world_size(group) == 1:
        return partial
    buf = partial.float()
    dist.all_reduce(buf, group=group)
    return buf.to(partial.dtype)


__all__ = ["init_tensor_parallel", "shard_rows", "shard_cols", "all_reduce_fp32", "Mapping", "TPCommunicator", "ParallelAttention",
           "get_parallel_groups"]
