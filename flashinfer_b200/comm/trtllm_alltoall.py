"""Legacy "prepare + comm" MoE all-to-all (reference flashinfer/comm/trtllm_alltoall.py: ``moe_comm_prepare_indices``,
``moe_local_gather``, ``moe_comm``, ``moe_prepare`` and the ``MnnvlMoe`` driver class; kernels
include/flashinfer/comm/trtllm_alltoall.cuh, csrc/trtllm_alltoall_prepare.cu).

The two-phase protocol is kept for engines that still drive it: index preparation is a handful of vectorised device ops on small
integer tensors (the routing of at most ``ep_size * max_tokens`` tokens), the data movement is the native indexed all-to-all-v of
csrc/comm/collectives.cu (``moe_comm_run``: rows pushed straight into the peer's staging slot over NVLink, then scattered into the
output; per-CTA cross-rank barriers with a watchdog).  New code should use the fused dispatch / combine of
:class:`flashinfer_b200.comm.MoeAlltoAll`.

Workspaces: the reference hands out ``MnnvlMemory`` tensors; here ``MnnvlMoe.get_moe_workspaces`` returns a ``[ep_size, words]``
uint64 view whose rows are backed by the symmetric heap of this process group - ``moe_comm`` recognises it by address.  A plain
CUDA tensor (single-rank use, as in the reference's single-GPU tests) is accepted for ``ep_size == 1``."""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

from .. import jit
from ..utils import stream_ptr
from .trtllm_moe_alltoall import MoeAlltoAll, moe_a2a_get_workspace_size_per_rank  # noqa: F401

_MAX_BLOCKS = 148
_max_sms = {"value": None}


def set_moe_max_usable_sm_count(max_sm_count: int) -> None:
    _max_sms["value"] = int(max_sm_count)


def get_moe_commworkspace_size_per_rank(ep_size: int) -> int:
    """Bytes of staging per rank (``FIB200_LEGACY_A2A_BYTES`` overrides; default 256 MiB: 2300 rows of 7168 bf16 per rank pair at
    EP 8)."""
    return int(os.environ.get("FIB200_LEGACY_A2A_BYTES", str(256 << 20)))


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


# ------------------------------------------------------------------ index preparation (device ops on small int tensors)
def moe_comm_prepare_indices(gathered_target_rank_ids: torch.Tensor, real_rank_token_count_cum_sum: Optional[torch.Tensor],
                             max_token_count_per_rank: int, expert_count: int, top_k: int, ep_rank: int, ep_size: int
                             ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """From the all-gathered ``[tokens of all ranks, top_k]`` target-rank table: what this rank sends (a token goes to a rank once,
    however many of its experts live there) and what it receives.

    Returns ``(local_gather_indices, send_rank_count_cumsum, send_rank_local_indices, recv_rank_count_cumsum,
    recv_rank_local_indices, backward_recv_rank_local_indices)``: received token ``i`` is row ``local_gather_indices[i]`` of the
    gathered tensors and lands in local slot ``recv_rank_local_indices[i]`` (source-rank major); send entry ``e`` ships local token
    ``send_rank_local_indices[e]``, and on the way back (combine) its result lands in row ``backward...[e] = token * top_k + k`` of
    the ``[tokens * top_k, hidden]`` buffer, ``k`` being the first expert slot of the token that points to that rank."""
    g = gathered_target_rank_ids.to(torch.int64)
    dev = g.device
    total = g.shape[0]
    if real_rank_token_count_cum_sum is not None:
        cum = real_rank_token_count_cum_sum.to(torch.int64)
        lo = int(cum[ep_rank - 1]) if ep_rank > 0 else 0
        hi = int(cum[ep_rank])
        src_rank = torch.searchsorted(cum, torch.arange(total, device=dev), right=True)
    else:
        lo, hi = ep_rank * max_token_count_per_rank, min((ep_rank + 1) * max_token_count_per_rank, total)
        src_rank = torch.arange(total, device=dev) // max_token_count_per_rank
    ranks = torch.arange(ep_size, device=dev)
    hit = (g.unsqueeze(-1) == ranks).any(1)                      # [total, ep_size]: token -> rank (deduplicated)
    mine = hit[lo:hi]                                            # my tokens
    send_pairs = mine.t().nonzero()                              # (rank, token) sorted by rank, then token
    send_rank_count_cumsum = mine.sum(0).cumsum(0).to(torch.int32)
    send_rank_local_indices = send_pairs[:, 1].to(torch.int32)
    first_k = (g[lo:hi].unsqueeze(-1) == ranks).int().argmax(1)  # [n_mine, ep_size]: first expert slot pointing at the rank
    backward = (send_pairs[:, 1] * top_k + first_k[send_pairs[:, 1], send_pairs[:, 0]]).to(torch.int32)
    recv_global = hit[:, ep_rank].nonzero()[:, 0]                # ascending global token index = source-rank major
    recv_counts = torch.bincount(src_rank[recv_global], minlength=ep_size)[:ep_size]
    recv_rank_count_cumsum = recv_counts.cumsum(0).to(torch.int32)
    recv_rank_local_indices = torch.arange(recv_global.numel(), device=dev, dtype=torch.int32)
    return (recv_global.to(torch.int32), send_rank_count_cumsum, send_rank_local_indices, recv_rank_count_cumsum,
            recv_rank_local_indices, backward)


def moe_local_gather(recv_rank_cum_sum: torch.Tensor, local_gather_indices: torch.Tensor, gathered_expert_ids: torch.Tensor,
                     gathered_scales: Optional[torch.Tensor], local_expert_ids: torch.Tensor, local_scales: Optional[torch.Tensor],
                     max_token_count_per_rank: int, expert_count: int, top_k: int, ep_rank: int, ep_size: int) -> None:
    """``local_*[i] = gathered_*[local_gather_indices[i]]`` for the received tokens, ``expert_count`` / 0 beyond them."""
    n = int(recv_rank_cum_sum[ep_size - 1])
    idx = local_gather_indices[:n].long()
    local_expert_ids.fill_(expert_count)
    local_expert_ids[:n] = gathered_expert_ids[idx].to(local_expert_ids.dtype)
    if local_scales is not None:
        local_scales.zero_()
        if gathered_scales is not None:
            local_scales[:n] = gathered_scales[idx].to(local_scales.dtype)


# ------------------------------------------------------------------ workspaces and the native data path
class _Ctx:
    """Staging region + signal pads of one expert-parallel group (symmetric heap; plain buffers for a single rank)."""

    def __init__(self, group, ep_size: int, nbytes: int, device: torch.device):
        self.ep_size, self.nbytes = ep_size, nbytes
        sig_bytes = 2 * _MAX_BLOCKS * 16 * 4
        if ep_size > 1:
            import torch.distributed as dist

            from .symm import SymmetricHeap

            self.group = group if group is not None else dist.group.WORLD
            self.heap = SymmetricHeap(self.group, sig_bytes + nbytes + 8192)
            _, sig_off = self.heap.alloc(sig_bytes)
            self.stage, stage_off = self.heap.alloc(nbytes)
            self.sig_tab = self.heap.peer_ptr_table(sig_off)
            self.buf_tab = self.heap.peer_ptr_table(stage_off)
            self.heap.barrier()
        else:
            self.sig = torch.zeros(sig_bytes, dtype=torch.uint8, device=device)
            self.stage = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self.sig_tab = torch.tensor([self.sig.data_ptr()], dtype=torch.int64)
            self.buf_tab = torch.tensor([self.stage.data_ptr()], dtype=torch.int64)
        self.epochs = torch.zeros(2 * _MAX_BLOCKS, dtype=torch.int32, device=device)
        self.token = self.stage.view(torch.int64)  # what the caller sees as "its" workspace row


_CTX: Dict[int, _Ctx] = {}          # workspace data_ptr -> context
_CTX_BY_SIZE: Dict[Tuple[int, int], _Ctx] = {}


def _ctx_for(all_workspaces: torch.Tensor, ep_size: int, device: torch.device, group=None) -> _Ctx:
    ctx = _CTX.get(all_workspaces.data_ptr())
    if ctx is not None:
        return ctx
    key = (ep_size, device.index if device.index is not None else torch.cuda.current_device())
    ctx = _CTX_BY_SIZE.get(key)
    if ctx is None:  # a workspace tensor we did not hand out (reference-style torch.zeros): bind a context of the default size
        ctx = _Ctx(group, ep_size, get_moe_commworkspace_size_per_rank(ep_size), device)
        _CTX_BY_SIZE[key] = ctx
    _CTX[all_workspaces.data_ptr()] = ctx
    return ctx


def moe_comm(input: torch.Tensor, send_rank_cum_sum: torch.Tensor, send_indices: torch.Tensor, output: torch.Tensor,
             recv_rank_cum_sum: torch.Tensor, recv_indices: torch.Tensor, all_workspaces: torch.Tensor, ep_rank: int,
             ep_size: int, group=None) -> None:
    """Indexed all-to-all-v: send entry ``e`` (rank ``j`` owns entries ``[send_cum[j - 1], send_cum[j])``) carries
    ``input[send_indices[e]]``; the ``k``-th row received from rank ``i`` is written to ``output[recv_indices[recv_cum[i - 1] + k]]``.
    Rows of ``output`` that receive nothing are left untouched."""
    if input.dim() != 2 or output.dim() != 2 or input.shape[1] != output.shape[1] or input.dtype != output.dtype:
        raise ValueError("moe_comm: input / output must be 2-D with the same row size and dtype")
    if not input.is_cuda:
        # CPU tensors: the oracle of the native kernel; several ranks go through torch.distributed (gloo) all_to_all_single
        n_send, n_recv = int(send_rank_cum_sum[ep_size - 1]), int(recv_rank_cum_sum[ep_size - 1])
        rows = input[send_indices[:n_send].long()].contiguous()
        if ep_size == 1:
            output[recv_indices[:n_recv].long()] = rows
            return
        import torch.distributed as dist

        sc, rc = send_rank_cum_sum.tolist(), recv_rank_cum_sum.tolist()
        in_split = [sc[0]] + [sc[i] - sc[i - 1] for i in range(1, ep_size)]
        out_split = [rc[0]] + [rc[i] - rc[i - 1] for i in range(1, ep_size)]
        recv = torch.empty(n_recv, input.shape[1], dtype=input.dtype)
        dist.all_to_all_single(recv, rows, output_split_sizes=out_split, input_split_sizes=in_split, group=group)
        output[recv_indices[:n_recv].long()] = recv
        return
    row_bytes = input.shape[1] * input.element_size()
    if row_bytes % 16 or input.stride(1) != 1 or output.stride(1) != 1:
        raise ValueError("moe_comm: rows must be contiguous multiples of 16 bytes")
    ctx = _ctx_for(all_workspaces, ep_size, input.device, group)
    cap_rows = ctx.nbytes // (ep_size * row_bytes)
    esz = input.element_size()
    jit.load("comm_collectives").call(
        "moe_comm_run", ctx.buf_tab, ctx.sig_tab, ctx.epochs, input, input.shape[0], input.stride(0) * esz,
        send_rank_cum_sum.to(torch.int32).contiguous(), send_indices.to(torch.int32).contiguous(), output, output.shape[0],
        output.stride(0) * esz, recv_rank_cum_sum.to(torch.int32).contiguous(), recv_indices.to(torch.int32).contiguous(), row_bytes,
        cap_rows, ep_rank, ep_size, _MAX_BLOCKS, stream_ptr(input))


def moe_prepare(experts_ids: torch.Tensor, scales: Optional[torch.Tensor], experts_statics: Optional[torch.Tensor],
                workspace: torch.Tensor, max_token_count_per_rank: int, ep_rank: int, ep_size: int, expert_count: int,
                slot_count: int, top_k: int, group=None):
    """Prepare phase without a user-side all-gather (reference :387): routing tables travel here (one all-gather of the
    ``[tokens, top_k]`` expert ids / scales over the group), then the same index preparation and local gather as above.

    Returns ``(prepared_local_experts, prepared_local_scales, send_cumsum, send_indices, recv_cumsum, recv_indices,
    backward_recv_indices, gathered_expert_statics)``."""
    dev = experts_ids.device
    n = experts_ids.shape[0]
    ids = torch.full((max_token_count_per_rank, top_k), slot_count, dtype=torch.int32, device=dev)
    ids[:n] = experts_ids.to(torch.int32)
    sc = torch.zeros(max_token_count_per_rank, top_k, dtype=torch.float32, device=dev)
    if scales is not None:
        sc[:n] = scales.float()
    cnt = torch.tensor([n], dtype=torch.int32, device=dev)
    if ep_size > 1:
        import torch.distributed as dist

        def gather(t):  # all_gather (list form: works on every backend, gloo included), rank-major concatenation
            parts = [torch.empty_like(t) for _ in range(ep_size)]
            dist.all_gather(parts, t.contiguous(), group=group)
            return parts

        g_ids, g_sc, g_cnt = torch.cat(gather(ids)), torch.cat(gather(sc)), torch.cat(gather(cnt))
        stat = torch.stack(gather(experts_statics)) if experts_statics is not None else None
    else:
        g_ids, g_sc, g_cnt = ids, sc, cnt
        stat = experts_statics[None] if experts_statics is not None else None
    per_rank = slot_count // ep_size
    target = torch.where(g_ids < slot_count, g_ids // per_rank, torch.full_like(g_ids, ep_size))  # padding rows target nobody
    (gather_idx, send_cum, send_idx, recv_cum, recv_idx, backward) = moe_comm_prepare_indices(
        target, None, max_token_count_per_rank, expert_count, top_k, ep_rank, ep_size)
    alloc = max_token_count_per_rank * ep_size
    local_ids = torch.empty(alloc, top_k, dtype=torch.int32, device=dev)
    local_sc = torch.empty(alloc, top_k, dtype=torch.float32, device=dev)
    moe_local_gather(recv_cum, gather_idx, g_ids, g_sc, local_ids, local_sc, max_token_count_per_rank, slot_count, top_k, ep_rank,
                     ep_size)
    return local_ids, (local_sc if scales is not None else None), send_cum, send_idx, recv_cum, recv_idx, backward, stat


class MnnvlMoe:
    """Driver class of the legacy protocol (reference :433): workspace accessors + the alltoallv / combine pair."""
    moe_workspace = None
    moe_prepare_workspace = None
    moe_workspace_tensor: Optional[torch.Tensor] = None
    moe_prepare_workspace_tensor: Optional[torch.Tensor] = None
    moe_mapping = None

    @staticmethod
    def get_moe_workspaces(mapping, config=None, group=None) -> torch.Tensor:
        if MnnvlMoe.moe_workspace_tensor is not None:
            return MnnvlMoe.moe_workspace_tensor
        ep = getattr(mapping, "tp_size", None) or getattr(mapping, "moe_ep_size", 1)
        dev = torch.device("cuda", torch.cuda.current_device())
        ctx = _Ctx(group, int(ep), get_moe_commworkspace_size_per_rank(int(ep)), dev)
        MnnvlMoe.moe_mapping, MnnvlMoe.moe_workspace = mapping, ctx
        MnnvlMoe.moe_workspace_tensor = ctx.token
        _CTX[ctx.token.data_ptr()] = ctx
        _CTX_BY_SIZE[(int(ep), dev.index)] = ctx
        return ctx.token

    @staticmethod
    def get_moe_prepare_workspace(mapping, config=None) -> torch.Tensor:
        if MnnvlMoe.moe_prepare_workspace_tensor is None:
            ep = getattr(mapping, "tp_size", None) or getattr(mapping, "moe_ep_size", 1)
            MnnvlMoe.moe_prepare_workspace_tensor = torch.zeros(get_moe_prepare_workspace_size_per_rank(int(ep)) // 8, dtype=torch.int64,
                                                                device="cuda")
        return MnnvlMoe.moe_prepare_workspace_tensor

    @staticmethod
    def compute_target_rank_id(token_selected_experts: torch.Tensor, expert_count: int, ep_size: int) -> torch.Tensor:
        assert expert_count % ep_size == 0, "expert_count should be divisible by ep_size"
        return token_selected_experts // (expert_count // ep_size)

    @staticmethod
    def mnnvl_moe_alltoallv_prepare_without_allgather(expert_ids, scales, expert_statics, workspace, max_token_count_per_rank: int,
                                                      ep_rank: int, ep_size: int, expert_count: int, slot_count: int, top_k: int,
                                                      group=None):
        (le, ls, send_cum, send_idx, recv_cum, recv_idx, backward, stat) = moe_prepare(
            expert_ids, scales, expert_statics, workspace, max_token_count_per_rank, ep_rank, ep_size, expert_count, slot_count, top_k,
            group=group)
        info = MoEAlltoallInfo(None, send_cum, send_idx, recv_cum, recv_idx, backward, max_token_count_per_rank * ep_size)
        return info, le, ls, stat

    @staticmethod
    def mnnvl_moe_alltoallv_prepare(gathered_target_rank_ids, real_rank_token_count_cumsum, gathered_expert_ids, gathered_scales,
                                    max_token_count_per_rank: int, expert_count: int, top_k: int, ep_rank: int, ep_size: int):
        (gather_idx, send_cum, send_idx, recv_cum, recv_idx, backward) = moe_comm_prepare_indices(
            gathered_target_rank_ids, real_rank_token_count_cumsum, max_token_count_per_rank, expert_count, top_k, ep_rank, ep_size)
        alloc = max_token_count_per_rank * ep_size
        dev = gathered_expert_ids.device
        local_expert_ids = torch.empty(alloc, top_k, dtype=torch.int32, device=dev)
        local_scales = torch.empty(alloc, top_k, dtype=torch.float32, device=dev)
        moe_local_gather(recv_cum, gather_idx, gathered_expert_ids, gathered_scales, local_expert_ids, local_scales,
                         max_token_count_per_rank, expert_count, top_k, ep_rank, ep_size)
        return (MoEAlltoallInfo(gather_idx, send_cum, send_idx, recv_cum, recv_idx, backward, alloc), local_expert_ids, local_scales)

    @staticmethod
    def mnnvl_moe_alltoallv(x: torch.Tensor, alltoall_info: MoEAlltoallInfo, workspace: torch.Tensor, ep_rank: int, ep_size: int,
                            group=None) -> torch.Tensor:
        assert x.dim() == 2, "only 2D tensor supported, please reshape."
        out = torch.empty(alltoall_info.local_token_allocation_count, x.shape[1], dtype=x.dtype, device=x.device)
        moe_comm(x, alltoall_info.send_rank_count_cumsum, alltoall_info.send_rank_local_indices, out,
                 alltoall_info.recv_rank_count_cumsum, alltoall_info.recv_rank_local_indices, workspace, ep_rank, ep_size, group=group)
        return out

    @staticmethod
    def mnnvl_moe_alltoallv_combine(x: torch.Tensor, alltoall_info: MoEAlltoallInfo, workspace: torch.Tensor, ep_rank: int,
                                    ep_size: int, top_k: int, token_count: int, group=None) -> torch.Tensor:
        assert x.dim() == 2, "2D tensor supported, please reshape."
        out = torch.zeros(token_count * top_k, x.shape[1], dtype=x.dtype, device=x.device)
        moe_comm(x, alltoall_info.recv_rank_count_cumsum, alltoall_info.recv_rank_local_indices, out,
                 alltoall_info.send_rank_count_cumsum, alltoall_info.backward_recv_rank_local_indices, workspace, ep_rank, ep_size,
                 group=group)
        return torch.sum(out.reshape(token_count, top_k, x.shape[1]), dim=1, keepdim=False)
