"""Decode context-parallel all-to-all (reference flashinfer/comm/dcp_alltoall.py:118-256, "Helix"): every rank holds partial attention
outputs and softmax statistics for all ``cp_size`` query slices and sends slice j to rank j - one kernel for both tensors (the
statistics ride behind the output rows), on the NVLink all-to-all of :class:`~flashinfer_b200.comm.collectives.NVLSCollectives`."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from .. import jit as _jit_acc
from .collectives import _MAX_BLOCKS, NVLSCollectives

_DCP: dict = {}


def decode_cp_a2a_workspace_size(cp_size: int) -> int:
    return 2 * (16 << 20) + 2 * _MAX_BLOCKS * 16 * 4 + 8192


def decode_cp_a2a_allocate_mnnvl_workspace(mapping, *, mnnvl_config=None, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Allocates the symmetric workspace for the CP group and returns an opaque handle tensor (int64 id)."""
    coll = NVLSCollectives(group, 16 << 20)
    h = torch.tensor([len(_DCP) + 1], dtype=torch.int64)
    _DCP[int(h)] = coll
    return h


def decode_cp_a2a_init_workspace(workspace: torch.Tensor, cp_rank: int, cp_size: int) -> None:
    """Nothing to reset (epoch barriers); kept for API compatibility.  Synchronises like the reference."""
    if torch.cuda.is_available():
        torch.cuda.current_stream().synchronize()


def decode_cp_a2a_alltoall(partial_o: torch.Tensor, softmax_stats: torch.Tensor, workspace: torch.Tensor, cp_rank: int,
                           cp_size: int, enable_pdl: Optional[bool] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """``partial_o [..., cp, D]`` and ``softmax_stats [..., cp, S]``: slice ``[..., j, :]`` goes to rank j.  Both tensors
    travel in ONE kernel launch (stats are packed behind the output rows)."""
    coll: NVLSCollectives = _DCP[int(workspace.reshape(-1)[0])]
    lead = partial_o.shape[:-2]
    rows = 1
    for d in lead:
        rows *= d
    D, S = partial_o.shape[-1], softmax_stats.shape[-1]
    ob = D * partial_o.element_size()
    sb = S * 4
    pad = (-(ob + sb)) % 16
    packed = torch.empty(rows, cp_size, ob + sb + pad, dtype=torch.uint8, device=partial_o.device)
    packed[..., :ob] = partial_o.reshape(rows, cp_size, D).contiguous().view(torch.uint8).view(rows, cp_size, ob)
    packed[..., ob:ob + sb] = softmax_stats.reshape(rows, cp_size, S).float().contiguous().view(torch.uint8).view(rows, cp_size, sb)
    recv = coll.all_to_all(packed)
    o = recv[..., :ob].contiguous().view(partial_o.dtype).view(*lead, cp_size, D)
    st = recv[..., ob:ob + sb].contiguous().view(torch.float32).view(*lead, cp_size, S)
    return o, st


get_dcp_alltoall_module = _jit_acc.module_accessor("comm_collectives")
