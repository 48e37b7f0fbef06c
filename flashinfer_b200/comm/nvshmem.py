"""NVSHMEM-style symmetric allocation / all-to-all (reference flashinfer/comm/nvshmem.py) on torch symmetric memory: this
image has no NVSHMEM, and on one NVSwitch domain symmetric memory + peer stores give the same programming model."""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.distributed as dist

_tensors = {}


def get_unique_id():
    return {"token": "torch-symmetric-memory"}


def alloc_empty_unique_id():
    return {}


def init(uid, rank: int, world_size: int) -> None:
    if not dist.is_initialized():
        raise RuntimeError("initialise torch.distributed first (the symmetric heap rendezvous uses its process group)")


def my_pe() -> int:
    return dist.get_rank()


def n_pes() -> int:
    return dist.get_world_size()


def malloc(shape: Sequence[int], dtype: torch.dtype, device: Optional[torch.device] = None) -> torch.Tensor:
    """Symmetric tensor: the same offset on every rank (collective call)."""
    import torch.distributed._symmetric_memory as symm_mem

    device = device or torch.device("cuda", torch.cuda.current_device())
    t = symm_mem.empty(*shape, dtype=dtype, device=device)
    _tensors[t.data_ptr()] = symm_mem.rendezvous(t, dist.group.WORLD)
    return t


def free_tensor(tensor: torch.Tensor) -> None:
    _tensors.pop(tensor.data_ptr(), None)


def alltoall(dest: torch.Tensor, source: torch.Tensor) -> None:
    """Equal-split all-to-all of ``source`` into ``dest`` (both ``[world * n, ...]``)."""
    if dist.get_backend() == "gloo":
        from ._p2p import all_to_all_uneven

        w = dist.get_world_size()
        all_to_all_uneven(list(dest.chunk(w)), list(source.chunk(w)))
    else:
        dist.all_to_all_single(dest, source)


def barrier_all() -> None:
    torch.cuda.synchronize() if torch.cuda.is_available() else None
    dist.barrier()


def barrier_all_on_current_stream() -> None:
    dist.barrier()


def finalize() -> None:
    _tensors.clear()
