"""Collectives behind one handler (reference flashinfer/comm/mixed_comm.py: ``MixedCommHandler``, ``run_mixed_comm(op, mode)``; ops AR /
AG / RS / AR+AG / RS+AR).  The reference mixes NVLS inside a node with NVSHMEM across nodes; on one NVSwitch domain every op runs on
:class:`~flashinfer_b200.comm.collectives.NVLSCollectives`."""
from __future__ import annotations

from enum import Enum
from typing import Optional

import torch
import torch.distributed as dist

from .. import jit as _jit_acc
from .collectives import NVLSCollectives

class MixedCommOp(Enum):
    ALLREDUCE = 0
    ALLGATHER = 1
    REDUCESCATTER = 2
    ALLREDUCE_ALLGATHER = 3
    REDUCESCATTER_ALLREDUCE = 4


class MixedCommMode(Enum):
    FUSED_NVLS = 0
    FUSED_P2P = 1
    NCCL = 2
    AUTOTUNE = 3


class MixedCommHandler:
    """TP x DP collectives on one NVSwitch domain.  ``local_tp_size * local_dp_size`` ranks form the group: AR / RS run
    inside each TP sub-group, AG across the DP sub-group (reference mixed_comm.py:143-421 topology model; the
    inter-node NVSHMEM legs do not exist on a single node)."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None, capacity_bytes: int = 64 << 20, hidden: int = 4096,
                 dtype: torch.dtype = torch.bfloat16, max_tokens: int = 8192, mode: MixedCommMode = MixedCommMode.FUSED_NVLS) -> None:
        from .allreduce import TPCommunicator

        self.group = group if group is not None else dist.group.WORLD
        self.mode = mode
        self.coll = NVLSCollectives(self.group, capacity_bytes, use_nvls=mode != MixedCommMode.FUSED_P2P)
        self.ar = TPCommunicator(self.group, max_tokens, hidden, dtype, use_nvls=mode != MixedCommMode.FUSED_P2P) \
            if self.coll._cuda else None

    def run(self, op: MixedCommOp, x: torch.Tensor) -> torch.Tensor:
        if self.mode == MixedCommMode.NCCL or not self.coll._cuda:
            return self._nccl(op, x)
        if op == MixedCommOp.ALLREDUCE:
            return self.ar.all_reduce(x)
        if op == MixedCommOp.ALLGATHER:
            return self.coll.all_gather(x)
        if op == MixedCommOp.REDUCESCATTER:
            return self.coll.reduce_scatter(x)
        if op == MixedCommOp.ALLREDUCE_ALLGATHER:
            return self.coll.all_gather(self.ar.all_reduce(x))
        if op == MixedCommOp.REDUCESCATTER_ALLREDUCE:
            return self.coll.reduce_scatter(x)
        raise ValueError(op)

    def _nccl(self, op: MixedCommOp, x: torch.Tensor) -> torch.Tensor:
        w = dist.get_world_size(self.group)
        if op == MixedCommOp.ALLREDUCE:
            y = x.clone()
            dist.all_reduce(y, group=self.group)
            return y
        if op in (MixedCommOp.ALLGATHER, MixedCommOp.ALLREDUCE_ALLGATHER):
            y = x.clone()
            if op == MixedCommOp.ALLREDUCE_ALLGATHER:
                dist.all_reduce(y, group=self.group)
            parts = [torch.empty_like(y) for _ in range(w)]
            dist.all_gather(parts, y, group=self.group)
            return torch.cat(parts, 0)
        y = x.clone()
        dist.all_reduce(y, group=self.group)
        n = x.shape[0] // w
        r = dist.get_rank(self.group)
        return y[r * n:(r + 1) * n].clone()


def run_mixed_comm(handler: MixedCommHandler, op: MixedCommOp, x: torch.Tensor) -> torch.Tensor:
    return handler.run(op, x)


get_mixed_comm_module = _jit_acc.module_accessor("comm_collectives")
