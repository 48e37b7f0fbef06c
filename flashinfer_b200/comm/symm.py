"""Symmetric heap over NVLink peer memory (+ NVLS multicast alias).

Parity: the reference's memory bootstrap layer — ``create_shared_buffer`` (CUDA IPC,
flashinfer/comm/cuda_ipc.py:197-237), ``_alloc_symm_buffer_bytes`` (torch symmetric memory,
flashinfer/comm/torch_symmetric_memory.py:45-76) and ``SymmDeviceMemory`` (VMM + multicast,
flashinfer/comm/mnnvl.py:876-1240).  Here there is ONE mechanism: a ``torch.distributed``
symmetric-memory allocation (VMM handles exchanged over the process group's store, peers mapped into
this process, plus a ``cuMulticast`` alias when the fabric supports NVLS).  Every comm kernel receives
``peer_ptrs`` (same offset on each rank), ``mc_ptr`` and a zero-initialised signal region.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


class SymmetricHeap:
    """One symmetric allocation per process group, carved with a bump allocator."""

    def __init__(self, group: Optional[dist.ProcessGroup], nbytes: int, device: Optional[torch.device] = None):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        nbytes = (nbytes + 4095) // 4096 * 4096
        self.nbytes = nbytes
        self.buffer = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.handle = symm_mem.rendezvous(self.buffer, self.group)
        self.buffer.zero_()
        self.peer_ptrs: List[int] = [int(p) for p in self.handle.buffer_ptrs]
        try:
            self.mc_ptr: int = int(self.handle.multicast_ptr) if self.handle.has_multicast_support() else 0
        except Exception:  # noqa: BLE001
            self.mc_ptr = 0
        self._off = 0
        torch.cuda.synchronize()
        dist.barrier(self.group)

    def alloc(self, nbytes: int, align: int = 1024):
        """Returns (local uint8 view, offset). Must be called in the same order on every rank."""
        off = (self._off + align - 1) // align * align
        if off + nbytes > self.nbytes:
            raise MemoryError(f"symmetric heap exhausted ({off + nbytes} > {self.nbytes})")
        self._off = off + nbytes
        return self.buffer[off : off + nbytes], off

    def peer_ptr_table(self, off: int) -> torch.Tensor:
        """Host int64 tensor with the address of ``off`` on every rank."""
        return torch.tensor([p + off for p in self.peer_ptrs], dtype=torch.int64)

    def mc(self, off: int) -> int:
        return self.mc_ptr + off if self.mc_ptr else 0

    def barrier(self) -> None:
        torch.cuda.synchronize()
        dist.barrier(self.group)
