"""The workspace object of the fused all-reduce APIs (reference flashinfer/comm/workspace_base.py).

The reference has three generations of workspaces (IPC buffers + Lamport flags, MNNVL multicast, vLLM signal buffers); here each of
them owns one :class:`~flashinfer_b200.comm.allreduce.TPCommunicator` (a symmetric heap with a multicast alias when the fabric
offers one), so reference call sites keep working."""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


class AllReduceFusionWorkspace:
    """Owns the symmetric heap + signal pads for one TP group."""

    backend = "nvls"

    def __init__(self, world_size: int, rank: int, max_token_num: int, hidden_dim: int, dtype: torch.dtype = torch.bfloat16,
                 group: Optional[dist.ProcessGroup] = None) -> None:
        self.world_size, self.rank = world_size, rank
        self.max_token_num, self.hidden_dim, self.dtype = max_token_num, hidden_dim, dtype
        from .allreduce import TPCommunicator

        self.comm = TPCommunicator(group, max_token_num, hidden_dim, dtype)
        self._destroyed = False

    def is_buffer_size_sufficient(self, tp_size: int, num_tokens: int, hidden_dim: int, dtype: Optional[torch.dtype] = None,
                                  use_oneshot=None, strategy=None) -> bool:
        """Argument order of the reference (workspace_base.py :54): ``(tp_size, num_tokens, hidden_dim, dtype, use_oneshot)``; the MNNVL
        workspace names its last argument ``strategy``.  One-shot / two-shot use the same heap here, so neither changes the answer."""
        if tp_size != self.world_size:
            return False
        esz = torch.empty(0, dtype=dtype or self.dtype).element_size()
        cap = self.max_token_num * self.hidden_dim * torch.empty(0, dtype=self.dtype).element_size()
        return num_tokens * hidden_dim * esz <= cap and hidden_dim == self.hidden_dim

    def destroy(self) -> None:
        self._destroyed = True
        self.comm = None

    @property
    def metadata(self) -> dict:
        return {"tp_size": self.world_size, "tp_rank": self.rank, "max_token_num": self.max_token_num,
                "hidden_dim": self.hidden_dim, "dtype": str(self.dtype)}
