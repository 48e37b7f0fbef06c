"""Module path of the reference (flashinfer/comm/nvshmem_allreduce.py): ``NVSHMEMAllReduce`` over the NVLS all-reduce kernel."""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


class NVSHMEMAllReduce:
    """Same constructor / ``all_reduce(inp, out)`` / ``shutdown()`` contract as the reference; the symmetric buffer comes from torch
    symmetric memory and the reduction is the in-switch ``multimem.ld_reduce`` kernel (csrc/comm/allreduce.cu)."""

    def __init__(self, local_rank: int, world_size: int, max_buffer_elements: int, dtype: torch.dtype, device: torch.device,
                 group: Optional[dist.ProcessGroup] = None, should_init: bool = True):
        from .allreduce import TPCommunicator

        self.local_rank, self.world_size, self.dtype, self.device, self.group = local_rank, world_size, dtype, device, group
        self.max_buffer_elements = max_buffer_elements
        hidden = 1024 if max_buffer_elements % 1024 == 0 else max_buffer_elements
        self._hidden = hidden
        self._comm = TPCommunicator(group, max(1, max_buffer_elements // hidden), hidden, dtype)

    def all_reduce(self, inp: torch.Tensor, out: torch.Tensor) -> None:
        n = inp.numel()
        if n % self._hidden or n > self.max_buffer_elements:
            raise ValueError("tensor does not fit the buffer geometry this NVSHMEMAllReduce was created with")
        res = self._comm.all_reduce(inp.reshape(-1, self._hidden))
        out.copy_(res.view_as(out))

    def shutdown(self) -> None:
        self._comm = None
