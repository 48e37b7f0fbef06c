"""vLLM-style custom all-reduce handle API (reference flashinfer/comm/vllm_ar.py:94-145; kernels include/flashinfer/comm/
vllm_custom_all_reduce.cuh).  A handle owns a :class:`~flashinfer_b200.comm.allreduce.TPCommunicator`; inputs are staged through its
symmetric heap, so buffers need no IPC registration and captured graphs replay without graph-buffer registration (epochs live in
device memory).  Exported both under the reference's short names (``init_custom_ar``, ``all_reduce`` ...) and with the ``vllm_`` prefix
``flashinfer.comm`` re-exports them with."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from .allreduce import TPCommunicator

_VLLM: dict = {}


def vllm_meta_size() -> int:
    return 0


def vllm_init_custom_ar(ipc_tensors=None, rank_data: Optional[torch.Tensor] = None, rank: int = 0, full_nvlink: bool = True,
                        group: Optional[dist.ProcessGroup] = None, max_size: int = 8 << 20, hidden: int = 4096,
                        dtype: torch.dtype = torch.bfloat16) -> int:
    """Returns an opaque handle.  Buffers do not need to be IPC-registered: inputs are staged into the symmetric heap."""
    esz = torch.empty(0, dtype=dtype).element_size()
    comm = TPCommunicator(group, max(1, max_size // (hidden * esz)), hidden, dtype)
    h = len(_VLLM) + 1
    _VLLM[h] = comm
    return h


def vllm_dispose(fa: int) -> None:
    _VLLM.pop(fa, None)


def vllm_all_reduce(fa: int, inp: torch.Tensor, out: torch.Tensor, reg_buffer: int = 0, reg_buffer_sz_bytes: int = 0,
                    num_ctas: int = 0) -> None:
    comm = _VLLM[fa]
    flat = inp.reshape(-1)
    if flat.numel() % comm.hidden == 0 and inp.dtype == comm.dtype:
        res = comm.allreduce_add_rmsnorm(flat.view(-1, comm.hidden), None, None)
        out.copy_(res.view(out.shape))
    else:
        out.copy_(comm.all_reduce(inp))


def vllm_register_buffer(fa: int, fake_ipc_ptrs: List[int]) -> None:
    """No-op: any tensor can be reduced (staged through the symmetric heap)."""


def vllm_register_graph_buffers(fa: int, handles: List[List[int]], offsets: List[List[int]]) -> None:
    """No-op: epochs live in device memory, so captured graphs replay without buffer registration."""


def vllm_get_graph_buffer_ipc_meta(fa: int) -> Tuple[List[int], List[int]]:
    return [], []


meta_size, init_custom_ar, dispose, all_reduce = vllm_meta_size, vllm_init_custom_ar, vllm_dispose, vllm_all_reduce
register_buffer, register_graph_buffers, get_graph_buffer_ipc_meta = vllm_register_buffer, vllm_register_graph_buffers, vllm_get_graph_buffer_ipc_meta
