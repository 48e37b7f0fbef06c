"""Legacy "prepare + comm" MoE all-to-all (reference flashinfer/comm/trtllm_alltoall.py, ``MnnvlMoe``).  The fused dispatch /
combine kernels of :mod:`flashinfer_b200.comm.moe_alltoall` replace the two-phase protocol; the helpers that are independent
of the wire protocol are provided, the phase-split entry points point to :class:`MoeAlltoAll`."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from .moe_alltoall import MoeAlltoAll, moe_a2a_get_workspace_size_per_rank  # noqa: F401

_max_sms = {"value": None}


def set_moe_max_usable_sm_count(max_sm_count: int) -> None:
    _max_sms["value"] = int(max_sm_count)


def get_moe_commworkspace_size_per_rank(ep_size: int) -> int:
    return moe_a2a_get_workspace_size_per_rank(ep_size, 8192, 8192 * 2)


def get_moe_prepare_workspace_size_per_rank(ep_size: int) -> int:
    return 4096 * ep_size


@dataclass
class MoEAlltoallInfo:
    local_gather_indices: Optional[torch.Tensor] = None
    send_rank_count_cumsum: Optional[torch.Tensor] = None
    send_rank_local_indices: Optional[torch.Tensor] = None
    recv_rank_count_cumsum: Optional[torch.Tensor] = None
    recv_rank_local_indices: Optional[torch.Tensor] = None
    backward_recv_rank_local_indices: Optional[torch.Tensor] = None
    local_token_allocation_count: int = 0


def _moved(name: str):
    def fn(*args, **kwargs):
        raise NotImplementedError(f"{name}: the two-phase prepare / comm protocol is replaced by the fused kernels of "
                                  "flashinfer_b200.comm.MoeAlltoAll (dispatch / combine); use that class")
    fn.__name__ = name
    return fn


moe_comm_prepare_indices = _moved("moe_comm_prepare_indices")
moe_local_gather = _moved("moe_local_gather")
moe_comm = _moved("moe_comm")
moe_prepare = _moved("moe_prepare")


class MnnvlMoe:
    """Name-compatible facade: workspace sizing works, the data path is :class:`MoeAlltoAll`."""
    moe_workspace = None
    moe_prepare_workspace = None

    @staticmethod
    def get_moe_workspaces(mapping, config=None):
        raise NotImplementedError("use flashinfer_b200.comm.MoeAlltoAll(mapping, ...), which owns its symmetric workspace")

    mnnvl_moe_alltoallv_prepare_without_allgather = staticmethod(_moved("mnnvl_moe_alltoallv_prepare_without_allgather"))
    mnnvl_moe_alltoallv = staticmethod(_moved("mnnvl_moe_alltoallv"))
    mnnvl_moe_alltoallv_combine = staticmethod(_moved("mnnvl_moe_alltoallv_combine"))
