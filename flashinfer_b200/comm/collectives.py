"""NVLS / NVLink collectives: all-gather, reduce-scatter, all-to-all (DCP) and the reference's "mixed_comm" front end.

Parity: reference flashinfer/comm/mixed_comm.py (MixedCommHandler / run_mixed_comm :423-1457, ops AR / AG / RS /
AR+AG / RS+AR), flashinfer/comm/dcp_alltoall.py:118-256 (decode_cp_a2a_*), flashinfer/comm/nvshmem*.py (symmetric
allocation + alltoall + barrier: covered by SymmetricHeap + these kernels).  Kernels: csrc/comm/collectives.cu.
CPU / gloo groups fall back to torch.distributed so that host logic stays testable.
"""
from __future__ import annotations

from enum import Enum
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from .. import jit
from ..utils import dtype_code, stream_ptr
from .allreduce import _ptr

_MAX_BLOCKS = 148


class NVLSCollectives:
    """Symmetric-heap backed collectives for one process group.

    ``capacity_bytes`` bounds the size of the *gathered* tensor (all-gather output / reduce-scatter input)."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None, capacity_bytes: int = 64 << 20, use_nvls: bool = True) -> None:
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.capacity = (capacity_bytes + 1023) // 1024 * 1024
        self._cuda = torch.cuda.is_available() and dist.get_backend(self.group) != "gloo"
        if not self._cuda:
            return
        from .symm import SymmetricHeap

        sig_bytes = 2 * _MAX_BLOCKS * 16 * 4
        self.heap = SymmetricHeap(self.group, sig_bytes + 2 * self.capacity + 8192)
        _, self._sig_off = self.heap.alloc(sig_bytes)
        self._regions = []
        for _ in range(2):  # ping-pong data regions
            v, off = self.heap.alloc(self.capacity)
            self._regions.append((v, off))
        self._sig_tab = self.heap.peer_ptr_table(self._sig_off)
        self._buf_tab = [self.heap.peer_ptr_table(off) for _, off in self._regions]
        self._epochs = torch.zeros(2 * _MAX_BLOCKS, dtype=torch.int32, device=self.heap.device)
        self.use_nvls = bool(use_nvls and self.heap.mc_ptr)
        self._turn = 0
        self._mod = jit.load("comm_collectives")
        self.heap.barrier()

    def _next(self):
        self._turn ^= 1
        v, off = self._regions[self._turn]
        return v, off, self._buf_tab[self._turn]

    # ---- all-gather
    def all_gather(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``x [n, ...]`` per rank -> ``[world * n, ...]`` (rank-major)."""
        if not self._cuda:
            parts = [torch.empty_like(x) for _ in range(self.world)]
            dist.all_gather(parts, x.contiguous(), group=self.group)
            res = torch.cat(parts, 0)
        else:
            xc = x.contiguous()
            nbytes = xc.numel() * xc.element_size()
            if nbytes % 16 or nbytes * self.world > self.capacity:
                raise ValueError("all_gather: shard must be a multiple of 16 bytes and fit the symmetric capacity")
            region, off, tab = self._next()
            self._mod.call("nvls_all_gather", tab, self._sig_tab, _ptr(self.heap.mc(off) if self.use_nvls else 0), self._epochs,
                           xc, 0, nbytes, self.rank, self.world, _MAX_BLOCKS, 1, stream_ptr(xc))
            res = region[: nbytes * self.world].view(x.dtype).view(self.world * x.shape[0], *x.shape[1:])
        if out is not None:
            out.copy_(res)
            return out
        return res

    def gathered_buffer(self, shard_shape: Tuple[int, ...], dtype: torch.dtype) -> torch.Tensor:
        """Symmetric input region for reduce_scatter (write partial results here to skip the staging copy)."""
        n = 1
        for d in shard_shape:
            n *= d
        nbytes = n * torch.empty(0, dtype=dtype).element_size() * self.world
        region, off, tab = self._next()
        self._pending = (off, tab)
        return region[:nbytes].view(dtype).view(self.world * shard_shape[0], *shard_shape[1:])

    # ---- reduce-scatter
    def reduce_scatter(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``x [world * n, ...]`` per rank -> sum over ranks of slice ``rank`` -> ``[n, ...]``."""
        n = x.shape[0] // self.world
        if not self._cuda:
            res = torch.empty(n, *x.shape[1:], dtype=x.dtype)
            dist.reduce_scatter_tensor(res, x.contiguous(), group=self.group) if dist.get_backend(self.group) != "gloo" else None
            if dist.get_backend(self.group) == "gloo":
                full = x.clone()
                dist.all_reduce(full, group=self.group)
                res = full[self.rank * n:(self.rank + 1) * n].clone()
        else:
            nbytes = x.numel() * x.element_size()
            if (nbytes // self.world) % 16 or nbytes > self.capacity:
                raise ValueError("reduce_scatter: shard must be a multiple of 16 bytes and fit the symmetric capacity")
            loc = None
            for (v, off), tab in zip(self._regions, self._buf_tab):
                if x.data_ptr() == v.data_ptr():
                    loc = (off, tab)
            if loc is None:
                region, off, tab = self._next()
                region[:nbytes].view(x.dtype).view(x.shape).copy_(x)
                loc = (off, tab)
            off, tab = loc
            res = out if out is not None else torch.empty(n, *x.shape[1:], dtype=x.dtype, device=x.device)
            self._mod.call("nvls_reduce_scatter", tab, self._sig_tab, _ptr(self.heap.mc(off) if self.use_nvls else 0), self._epochs,
                           res, 0, nbytes // self.world, dtype_code(x.dtype), self.rank, self.world, _MAX_BLOCKS, 1, stream_ptr(x))
            return res
        if out is not None:
            out.copy_(res)
            return out
        return res

    # ---- all-to-all along a dim of size world
    def all_to_all(self, x: torch.Tensor) -> torch.Tensor:
        """``x [rows, world, ...]``: slice ``[:, j]`` goes to rank j; returns ``[rows, world, ...]`` with ``[:, i]`` from rank i."""
        rows = x.shape[0]
        if x.shape[1] != self.world:
            raise ValueError("all_to_all: dim 1 must equal the group size")
        if not self._cuda:
            send = x.transpose(0, 1).contiguous()
            recv = torch.empty_like(send)
            dist.all_to_all_single(recv, send, group=self.group)
            return recv.transpose(0, 1).contiguous()
        xc = x.contiguous()
        row_bytes = xc[0, 0].numel() * xc.element_size()
        nbytes = xc.numel() * xc.element_size()
        if row_bytes % 16 or nbytes > self.capacity:
            raise ValueError("all_to_all: per-peer rows must be multiples of 16 bytes and fit the symmetric capacity")
        region, off, tab = self._next()
        self._mod.call("p2p_all_to_all", tab, self._sig_tab, self._epochs, xc, 0, rows, row_bytes, self.rank, self.world,
                       _MAX_BLOCKS, 1, stream_ptr(xc))
        return region[:nbytes].view(x.dtype).view(x.shape)
