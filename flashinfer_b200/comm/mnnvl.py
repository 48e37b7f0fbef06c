"""Multi-node-NVLink memory helpers (reference flashinfer/comm/mnnvl.py).  The reference builds its own VMM + fabric-handle
exchange; here the same objects sit on torch symmetric memory (VMM allocations with a multicast alias when the switch
supports NVLS), which is what every kernel in ``csrc/comm`` addresses."""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.distributed as dist


def round_up(val: int, gran: int) -> int:
    return (val + gran - 1) // gran * gran


class CommBackend(ABC):
    @abstractmethod
    def Get_rank(self) -> int: ...

    @abstractmethod
    def Get_size(self) -> int: ...

    @abstractmethod
    def allgather(self, data): ...

    @abstractmethod
    def barrier(self) -> None: ...


class TorchDistBackend(CommBackend):
    """``torch.distributed`` process group behind the reference's MPI-like interface."""

    def __init__(self, group=None):
        self.group = group if group is not None else dist.group.WORLD

    def Get_rank(self) -> int:
        return dist.get_rank(self.group)

    def Get_size(self) -> int:
        return dist.get_world_size(self.group)

    def allgather(self, data):
        out = [None] * self.Get_size()
        dist.all_gather_object(out, data, group=self.group)
        return out

    def bcast(self, data, root: int = 0):
        box = [data]
        dist.broadcast_object_list(box, src=root, group=self.group)
        return box[0]

    def barrier(self) -> None:
        dist.barrier(self.group)

    def Split(self, color: int, key: int) -> "TorchDistBackend":
        return self


@dataclass
class MnnvlConfig:
    comm_backend: Optional[CommBackend] = None
    allocation_granularity: int = 0
    fabric_page_size: int = 1 << 29


def is_mnnvl_fabric_supported(device_idx: int) -> bool:
    """Peer-addressable symmetric memory is available on every NVSwitch box (intra-node); fabric handles are not used."""
    return torch.cuda.is_available()


class SymmDeviceMemory:
    """A symmetric buffer of ``buf_size`` bytes per rank with unicast pointers to every peer and (if supported) a multicast alias."""

    def __init__(self, buf_size: int, group_size: Optional[int] = None, group_rank: Optional[int] = None, device_idx: Optional[int] = None,
                 is_multi_node: bool = False, comm_backend_for_handle_transfer: Optional[CommBackend] = None,
                 enable_multicast: bool = True, group=None):
        from .symm import SymmetricHeap

        self.heap = SymmetricHeap(group, buf_size)
        self.buf_size = buf_size
        self.uc_ptrs: List[int] = list(self.heap.peer_ptrs)
        self.mc_ptr: int = self.heap.mc_ptr if enable_multicast else 0
        self.group_rank, self.group_size = self.heap.rank, self.heap.world

    def get_multicast_ptr(self) -> int:
        return self.mc_ptr

    def get_unicast_ptr(self, rank: int) -> int:
        return self.uc_ptrs[rank]

    def get_signal_pad_ptrs_host(self) -> List[int]:
        return self.uc_ptrs

    def get_buffer_ptrs_host(self) -> List[int]:
        return self.uc_ptrs

    def get_world_size(self) -> int:
        return self.group_size

    def get_rank(self) -> int:
        return self.group_rank

    def get_usable_buffer_size(self) -> int:
        return self.buf_size


class McastGPUBuffer:
    """Multicast-addressable buffer (reference mnnvl.py:1297)."""

    def __init__(self, buf_size: int, group_size: Optional[int] = None, group_rank: Optional[int] = None,
                 device: Optional[torch.device] = None, mn_nvlink: bool = True, comm_backend_for_handle_transfer=None, group=None):
        self.mcast_device_memory = SymmDeviceMemory(buf_size, group_size, group_rank, None, mn_nvlink, comm_backend_for_handle_transfer,
                                                    True, group)
        self.buf_size = buf_size
        self.local_device = device or torch.device("cuda", torch.cuda.current_device())

    def lamport_initialize(self, rank: int, dtype: torch.dtype) -> None:
        """Fill the local buffer with the negative-zero sentinel of the reference's Lamport protocol."""
        buf = self.mcast_device_memory.heap.buffer
        if dtype in (torch.float16, torch.bfloat16):
            buf.view(torch.int16).fill_(-32768)
        else:
            buf.view(torch.int32).fill_(-2147483648)
        torch.cuda.synchronize()

    def get_multicast_ptr(self) -> int:
        return self.mcast_device_memory.get_multicast_ptr()

    def get_unicast_ptr(self, rank: int) -> int:
        return self.mcast_device_memory.get_unicast_ptr(rank)

    def get_buffer_ptrs_dev(self) -> torch.Tensor:
        return torch.tensor(self.mcast_device_memory.uc_ptrs, dtype=torch.int64, device=self.local_device)


class MnnvlMemory:
    """A strided view over every rank's slice of one symmetric allocation (reference mnnvl.py:326)."""

    def __init__(self, mapping, size: int, group=None):
        from .symm import SymmetricHeap

        self.mapping, self.segment_size = mapping, size
        self.heap = SymmetricHeap(group, size)
        self.rank, self.world = self.heap.rank, self.heap.world

    def as_torch_strided_tensor(self, dtype: torch.dtype) -> torch.Tensor:
        """``[world, size / itemsize]``: row ``r`` is rank ``r``'s slice (peer memory, addressable from this rank)."""
        from .dlpack_utils import pack_strided_memory

        esz = torch.empty(0, dtype=dtype).element_size()
        rows = [pack_strided_memory(p, self.segment_size, self.segment_size, 1, dtype, self.heap.device)[0] for p in self.heap.peer_ptrs]
        # peers live at unrelated virtual addresses: expose them as a list-backed stack (a view per peer)
        return _PeerRows(rows, self.segment_size // esz)

    @staticmethod
    def supports_mnnvl() -> bool:
        return torch.cuda.is_available()

    @staticmethod
    def initialize() -> None:
        return None

    @staticmethod
    def set_comm_from_config(mapping, config: Optional[MnnvlConfig] = None) -> None:
        return None


class _PeerRows:
    """Indexable ``[world][n]`` collection of peer views (each row is a real tensor over that peer's memory)."""

    def __init__(self, rows, n):
        self.rows, self.shape = rows, (len(rows), n)

    def __getitem__(self, i):
        return self.rows[i]

    def __len__(self):
        return len(self.rows)
