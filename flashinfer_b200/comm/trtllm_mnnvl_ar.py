"""MNNVL all-reduce entry points (reference flashinfer/comm/trtllm_mnnvl_ar.py).  On one NVSwitch domain "multi-node NVLink"
and NVLS are the same fabric: these calls run the in-switch (``multimem``) fused all-reduce kernels of ``csrc/comm/allreduce.cu``."""
from __future__ import annotations

from enum import Enum
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from .trtllm_ar import AllReduceFusionPattern, allreduce_fusion  # noqa: F401
from .workspace_base import AllReduceFusionWorkspace


class MNNVLAllReduceFusionWorkspace(AllReduceFusionWorkspace):
    backend = "mnnvl"


class MNNVLAllreduceFusionStrategy(Enum):
    ONESHOT = 0
    TWOSHOT = 1
    AUTO = 99

    @staticmethod
    def select_strategy(tp_size: int, num_tokens: int, hidden_dim: int, dtype: torch.dtype) -> "MNNVLAllreduceFusionStrategy":
        nbytes = num_tokens * hidden_dim * torch.empty(0, dtype=dtype).element_size()
        return MNNVLAllreduceFusionStrategy.ONESHOT if nbytes <= (1 << 20) else MNNVLAllreduceFusionStrategy.TWOSHOT


def mpi_barrier() -> None:
    if dist.is_initialized():
        dist.barrier()


def get_allreduce_mnnvl_workspace(mapping, dtype: torch.dtype, comm_backend_for_handle_transfer=None, buffer_size_in_bytes=None,
                                  max_token_num: int = 8192, hidden_dim: int = 8192, group=None) -> MNNVLAllReduceFusionWorkspace:
    g = group if group is not None else dist.group.WORLD
    return MNNVLAllReduceFusionWorkspace(dist.get_world_size(g), dist.get_rank(g), max_token_num, hidden_dim, dtype, g)


def trtllm_mnnvl_allreduce(input: torch.Tensor, workspace: MNNVLAllReduceFusionWorkspace, launch_with_pdl: bool,
                           output: Optional[torch.Tensor] = None,
                           strategy: MNNVLAllreduceFusionStrategy = MNNVLAllreduceFusionStrategy.AUTO) -> torch.Tensor:
    """Sum of ``input [tokens, hidden]`` over the group."""
    return allreduce_fusion(input, workspace, AllReduceFusionPattern.kAllReduce, launch_with_pdl, output=output)


def trtllm_mnnvl_fused_allreduce_add_rmsnorm(input: torch.Tensor, residual_in: torch.Tensor, gamma: torch.Tensor,
                                             workspace: MNNVLAllReduceFusionWorkspace, epsilon: Optional[float] = None,
                                             output: Optional[torch.Tensor] = None, residual_out: Optional[torch.Tensor] = None,
                                             launch_with_pdl: bool = False,
                                             strategy: MNNVLAllreduceFusionStrategy = MNNVLAllreduceFusionStrategy.AUTO
                                             ) -> Tuple[torch.Tensor, torch.Tensor]:
    """``residual_out = allreduce(input) + residual_in``; ``output = rmsnorm(residual_out) * gamma``.  Returns both."""
    eps = float(epsilon) if epsilon is not None else float(torch.finfo(input.dtype).eps)
    res = residual_out if residual_out is not None else torch.empty_like(residual_in)
    norm = allreduce_fusion(input, workspace, AllReduceFusionPattern.kARResidualRMSNorm, launch_with_pdl, residual_in=residual_in,
                            residual_out=res, norm_out=output, rms_gamma=gamma, rms_eps=eps)
    return norm, res


# legacy names of the reference
trtllm_mnnvl_all_reduce = trtllm_mnnvl_allreduce
trtllm_mnnvl_fused_allreduce_rmsnorm = trtllm_mnnvl_fused_allreduce_add_rmsnorm
