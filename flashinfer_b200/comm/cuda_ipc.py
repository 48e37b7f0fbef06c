"""Peer-addressable buffers and the CUDA runtime calls around them (reference flashinfer/comm/cuda_ipc.py: ``create_shared_buffer``
:197-237, ``free_shared_buffer``, ``CudaRTLibrary`` :70).  The reference exchanges ``cudaIpcMemHandle_t`` over torch.distributed; here a
shared buffer is an allocation of a symmetric-memory heap (same peer-pointer arithmetic as the rest of ``comm``)."""
from __future__ import annotations

from typing import Optional

import torch

_SHARED_HEAPS: dict = {}


def create_shared_buffer(size_in_bytes: int, group=None):
    """Reference cuda_ipc.py:197: a buffer that every rank of ``group`` can address; returns the list of per-rank device
    pointers (index = rank).  Backed by a symmetric-memory heap here (same address arithmetic as the rest of ``comm``)."""
    from .symm import SymmetricHeap

    heap = SymmetricHeap(group, int(size_in_bytes) + 1024)
    _, off = heap.alloc(int(size_in_bytes))
    ptrs = [int(p) for p in heap.peer_ptr_table(off).tolist()]
    _SHARED_HEAPS[ptrs[heap.rank]] = heap
    heap.barrier()
    return ptrs


def free_shared_buffer(pointers, group=None) -> None:
    """Drop the heap behind a :func:`create_shared_buffer` result (the memory is released with the last reference)."""
    for p in pointers:
        _SHARED_HEAPS.pop(int(p), None)


class CudaRTLibrary:
    """Minimal ctypes view of libcudart with the calls the reference's IPC helpers use (cuda_ipc.py:70)."""

    def __init__(self, so_file: Optional[str] = None):
        import ctypes
        import glob
        import os

        if so_file is None:
            cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + \
                glob.glob("/usr/local/cuda/lib64/libcudart.so*")
            so_file = cands[0] if cands else "libcudart.so"
        self.lib = ctypes.CDLL(so_file)
        self._ct = ctypes

    def _chk(self, rc: int) -> None:
        if rc != 0:
            raise RuntimeError(f"CUDART error {rc}")

    def cudaSetDevice(self, device: int) -> None:
        self._chk(self.lib.cudaSetDevice(device))

    def cudaDeviceSynchronize(self) -> None:
        self._chk(self.lib.cudaDeviceSynchronize())

    def cudaMalloc(self, size: int):
        p = self._ct.c_void_p()
        self._chk(self.lib.cudaMalloc(self._ct.byref(p), self._ct.c_size_t(size)))
        return p

    def cudaFree(self, p) -> None:
        self._chk(self.lib.cudaFree(p))

    def cudaMemset(self, p, value: int, count: int) -> None:
        self._chk(self.lib.cudaMemset(p, value, self._ct.c_size_t(count)))
