"""Fused all-gather + matmul (sequence-parallel column-parallel linear) in ONE kernel per rank.

Parity: reference flashinfer/comm/all_gather_matmul/all_gather_matmul.py:52-75 (``all_gather_matmul(inp, w, group)``).
Kernel: csrc/gemm/gemm_allgather_sm100.cu.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .. import jit
from ..utils import dtype_code, stream_ptr
from .allreduce import _ptr

_MAX_ROW_TILES = 1024
_NUM_SMS = 148


class AllGatherMatmul:
    """``out [world*Ml, N] = all_gather(x [Ml, K]) @ w[N, K].T``; ``Ml`` must be a multiple of 128."""

    def __init__(self, group: Optional[dist.ProcessGroup], max_rows_per_rank: int, k: int, dtype: torch.dtype = torch.bfloat16,
                 use_nvls: bool = True) -> None:
        from .symm import SymmetricHeap

        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.max_ml, self.k, self.dtype = max_rows_per_rank, k, dtype
        esz = torch.empty(0, dtype=dtype).element_size()
        gbytes = (self.world * max_rows_per_rank * k * esz + 1023) // 1024 * 1024
        fbytes = _MAX_ROW_TILES * 4
        self.heap = SymmetricHeap(self.group, 2 * gbytes + fbytes + 16384)
        self._gath = []
        for _ in range(2):
            v, off = self.heap.alloc(gbytes)
            self._gath.append((v, off, self.heap.peer_ptr_table(off)))
        _, self._flag_off = self.heap.alloc(fbytes)
        self._flag_tab = self.heap.peer_ptr_table(self._flag_off)
        self._expect = torch.zeros(_NUM_SMS * _MAX_ROW_TILES, dtype=torch.int32, device=self.heap.device)
        self.use_nvls = bool(use_nvls and self.heap.mc_ptr)
        self._turn = 0
        self._mod = jit.load("gemm_allgather_sm100")
        self.heap.barrier()

    def __call__(self, x: torch.Tensor, w: torch.Tensor, out: Optional[torch.Tensor] = None, return_gathered: bool = False,
                 slices: int = 4, bn: int = 0):
        Ml, K = x.shape
        N = w.shape[0]
        if K != self.k or Ml > self.max_ml or x.dtype != self.dtype:
            raise ValueError("AllGatherMatmul: shape / dtype does not match the communicator")
        if Ml % 128:
            raise ValueError("AllGatherMatmul: rows per rank must be a multiple of 128")
        if x.stride(1) != 1:
            x = x.contiguous()
        if w.stride(1) != 1:
            w = w.contiguous()
        self._turn ^= 1
        g, goff, gtab = self._gath[self._turn]
        M = Ml * self.world
        gathered = g[: M * K * x.element_size()].view(x.dtype).view(M, K)
        if out is None:
            out = torch.empty(M, N, dtype=x.dtype, device=x.device)
        mc = self.heap.mc if self.use_nvls else (lambda off: 0)
        self._mod.call("allgather_gemm_nt", x, w, gathered, out, Ml, N, K, x.stride(0), w.stride(0), out.stride(0),
                       dtype_code(x.dtype), gtab, self._flag_tab, _ptr(mc(goff)), _ptr(mc(self._flag_off)), self._expect,
                       _MAX_ROW_TILES, self.rank, self.world, slices, bn, 1, stream_ptr(x))
        return (out, gathered) if return_gathered else out


_CACHE: dict = {}


def all_gather_matmul(inp: torch.Tensor, w: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``inp [Ml, K]`` (row shard), ``w [N, K]`` -> ``[world * Ml, N]``."""
    g = group if group is not None else dist.group.WORLD
    key = (id(g), inp.shape[1], inp.dtype)
    comm = _CACHE.get(key)
    if comm is None or comm.max_ml < inp.shape[0]:
        comm = AllGatherMatmul(g, max(inp.shape[0], 2048), inp.shape[1], inp.dtype)
        _CACHE[key] = comm
    return comm(inp, w, out)
