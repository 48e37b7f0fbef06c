"""TRT-LLM style fused all-reduce entry points (reference flashinfer/comm/trtllm_ar.py: enums :37-113, workspace creation :430-760,
trtllm_custom_all_reduce :809, trtllm_allreduce_fusion :951, trtllm_moe_allreduce_fusion :1062, trtllm_moe_finalize_allreduce_fusion
:1140) on the kernels of csrc/comm/allreduce.cu: the one-shot push all-reduce with fused prologues (MoE reduction / finalize) and
epilogues (residual add, RMSNorm, fp8 / NVFP4 quantisation with any scale-factor layout) for decode-sized messages, the in-switch
``multimem.ld_reduce`` kernel for large ones."""
from __future__ import annotations

from typing import Optional, Union

import torch
import torch.distributed as dist

from .workspace_base import AllReduceFusionWorkspace


class AllReduceStrategyType:
    NCCL = 0
    MIN_LATENCY = 1
    UB = 2
    AUTO = 3
    ONESHOT = 4
    TWOSHOT = 5
    LOWPRECISION = 6
    MNNVL = 7


class AllReduceStrategyConfig:
    USE_MEMCPY = 1 << 0
    PUSH_MODE = 1 << 1


class AllReduceFusionOp:
    NONE = 0
    RESIDUAL_RMS_NORM = 1
    LAST_PROCESS_FOR_UB = 2
    RESIDUAL_RMS_PREPOST_NORM = 3
    RESIDUAL_RMS_NORM_QUANT_FP8 = 4
    RESIDUAL_RMS_NORM_QUANT_NVFP4 = 5
    RESIDUAL_RMS_NORM_OUT_QUANT_FP8 = 6
    RESIDUAL_RMS_NORM_OUT_QUANT_NVFP4 = 7
    MOE_ALLREDUCE_RESIDUAL_RMS_NORM = 8
    MOE_FINALIZE_ALLREDUCE_RESIDUAL_RMS_NORM = 9


class AllReduceFusionPattern:
    kAllReduce = 0
    kARResidualRMSNorm = 1
    kARResidualRMSNormFP8Quant = 2
    kARResidualRMSNormFP4Quant = 3
    kARResidualRMSNormOutFP8Quant = 4
    kARResidualRMSNormOutFP4Quant = 5
    kMoEReductionARResidualRMSNorm = 6
    kMoEFinalizeARResidualRMSNorm = 7
    kARResidualRMSNormPerTokenGroupFP8PackedQuant = 8
    kARResidualRMSNormOutPerTokenGroupFP8PackedQuant = 9


class QuantizationSFLayout:
    SWIZZLED_128x4 = 0
    SWIZZLED_8x4 = 1
    LINEAR = 2


def compute_fp4_swizzled_layout_sf_size(total_row: int, total_column: int) -> int:
    return (total_row + 127) // 128 * 128 * ((total_column + 3) // 4 * 4)


class TRTLLMAllReduceFusionWorkspace(AllReduceFusionWorkspace):
    backend = "trtllm"


def trtllm_create_ipc_workspace_for_all_reduce_fusion(tp_rank: int, tp_size: int, max_token_num: int, hidden_dim: int,
                                                      use_fp32_lamport: bool = False, group: Optional[dist.ProcessGroup] = None,
                                                      create_metadata: bool = False, comm_backend=None, use_symm_dev_mem: bool = False,
                                                      *, dtype: torch.dtype = torch.bfloat16):
    """Returns ``(handles, workspace)`` like the reference; ``workspace`` is the object to pass as ``workspace_ptrs``.
    ``comm_backend`` (how the reference exchanges IPC handles) and ``use_symm_dev_mem`` (its allocator choice) have no counterpart:
    the heap is always a symmetric allocation of this library."""
    ws = TRTLLMAllReduceFusionWorkspace(tp_size, tp_rank, max_token_num, hidden_dim, dtype, group)
    if create_metadata:
        return [ws], ws, ws.metadata
    return [ws], ws


def trtllm_destroy_ipc_workspace_for_all_reduce_fusion(workspace, group=None) -> None:
    for w in (workspace if isinstance(workspace, (list, tuple)) else [workspace]):
        w.destroy()


# ------------------------------------------------------------------ legacy pointer-table workspace (reference trtllm_ar.py :430, :809)
_LEGACY: dict = {}          # handle value -> workspace


class _IpcHandles(list):
    """What ``trtllm_create_ipc_workspace_for_all_reduce`` returns: seven rows of ``tp_size`` integers, shaped like the reference's
    peer pointer tables (comm buffers x2, barrier flags x2, Lamport buffers x3).  The integers are opaque handles of this library's
    symmetric heap (NOT device addresses): ``trtllm_custom_all_reduce`` resolves the communicator from any of them."""

    workspace = None


def trtllm_create_ipc_workspace_for_all_reduce(rank: int, tp_size: int, max_token_num: int, hidden_dim: int,
                                               group: Optional[dist.ProcessGroup] = None, *, dtype: torch.dtype = torch.bfloat16):
    ws = TRTLLMAllReduceFusionWorkspace(tp_size, rank, max_token_num, hidden_dim, dtype, group)
    base = (id(ws) & 0xFFFFFFFFFF) << 16
    handles = _IpcHandles([base + row * 256 + r for r in range(tp_size)] for row in range(7))
    handles.workspace = ws
    for row in handles:
        for h in row:
            _LEGACY[h] = ws
    return handles


def trtllm_destroy_ipc_workspace_for_all_reduce(workspace, group=None) -> None:
    seen = set()
    for row in workspace:
        for h in (row if isinstance(row, (list, tuple)) else [row]):
            ws = _LEGACY.pop(h, None) if isinstance(h, int) else h
            if ws is not None and id(ws) not in seen:
                seen.add(id(ws))
                ws.destroy()


def trtllm_lamport_initialize(buffer_ptr: int, size: int, dtype: torch.dtype) -> None:
    """No-op for callers: the one-shot push kernel keeps its own sentinel-initialised rotating buffers and re-arms them itself."""


def trtllm_lamport_initialize_all(buffer_0_ptr: int, buffer_1_ptr: int, buffer_2_ptr: int, size: int, dtype: torch.dtype) -> None:
    """No-op (see :func:`trtllm_lamport_initialize`)."""


# ------------------------------------------------------------------ fused all-reduce entry points
def _quant_after(norm: torch.Tensor, pattern: int, quant_out, scale_out, scale_factor, layout_code):
    P = AllReduceFusionPattern
    if pattern in (P.kARResidualRMSNormFP8Quant, P.kARResidualRMSNormOutFP8Quant):
        s = scale_factor if isinstance(scale_factor, torch.Tensor) else torch.tensor(float(scale_factor or 1.0), device=norm.device)
        q = (norm.float() / s.float()).clamp(-448, 448).to(torch.float8_e4m3fn)
        if quant_out is not None:
            quant_out.copy_(q.view(quant_out.shape))
        return q
    if pattern in (P.kARResidualRMSNormFP4Quant, P.kARResidualRMSNormOutFP4Quant):
        from ..quantization.fp4 import SfLayout, nvfp4_quantize

        gs = scale_factor if isinstance(scale_factor, torch.Tensor) else torch.tensor(float(scale_factor or 1.0), device=norm.device)
        lay = SfLayout.layout_linear if layout_code == QuantizationSFLayout.LINEAR else SfLayout.layout_128x4
        q, sf = nvfp4_quantize(norm, gs, sfLayout=lay)
        if quant_out is not None:
            quant_out.view(torch.uint8).reshape(-1)[: q.numel()].copy_(q.view(torch.uint8).reshape(-1))
        if scale_out is not None:
            scale_out.view(torch.uint8).reshape(-1)[: sf.numel()].copy_(sf.view(torch.uint8).reshape(-1))
        return q
    return None


_SF_LAYOUT_NAME = {QuantizationSFLayout.SWIZZLED_128x4: "128x4", QuantizationSFLayout.SWIZZLED_8x4: "8x4", QuantizationSFLayout.LINEAR: "linear"}


def _push_fused(comm: TPCommunicator, x, pattern: int, *, allreduce_out=None, residual_in=None, residual_out=None, rms_gamma=None,
                rms_eps=1e-6, norm_out=None, quant_out=None, scale_out=None, scale_factor=None, layout_code=None, pdl=True,
                moe_reduction=None, moe_finalize=None) -> None:
    """Every pattern of the reference in ONE kernel (one-shot push all-reduce with fused prologue / epilogue)."""
    P = AllReduceFusionPattern
    quant = "none"
    if pattern in (P.kARResidualRMSNormFP8Quant, P.kARResidualRMSNormOutFP8Quant):
        quant = "fp8"
    elif pattern in (P.kARResidualRMSNormFP4Quant, P.kARResidualRMSNormOutFP4Quant) or (quant_out is not None and pattern in (
            P.kMoEReductionARResidualRMSNorm, P.kMoEFinalizeARResidualRMSNorm)):
        quant = "nvfp4"
    comm.push_allreduce(x, ar_out=allreduce_out, residual_in=residual_in, residual_out=residual_out, rms_gamma=rms_gamma,
                        norm_out=norm_out, eps=rms_eps, quant=quant, quant_out=quant_out, scale_out=scale_out, scale_factor=scale_factor,
                        sf_layout=_SF_LAYOUT_NAME.get(layout_code, "128x4"), moe_reduction=moe_reduction, moe_finalize=moe_finalize,
                        enable_pdl=pdl)


def allreduce_fusion(input: torch.Tensor, workspace: AllReduceFusionWorkspace, pattern: int, launch_with_pdl: bool = False,
                     trigger_completion_at_end: bool = True, output: Optional[torch.Tensor] = None,
                     residual_out: Optional[torch.Tensor] = None, norm_out: Optional[torch.Tensor] = None,
                     quant_out: Optional[torch.Tensor] = None, scale_out: Optional[torch.Tensor] = None,
                     residual_in: Optional[torch.Tensor] = None, rms_gamma: Optional[torch.Tensor] = None, rms_eps: float = 1e-6,
                     scale_factor: Optional[Union[torch.Tensor, float]] = None, layout_code: Optional[int] = None,
                     use_oneshot: Optional[bool] = None, fp32_acc: bool = False, **moe_kwargs) -> torch.Tensor:
    """Unified fused all-reduce (patterns 0-5): sum over ranks [+ residual add + RMSNorm [+ fp8 / nvfp4 quant]].

    Up to ``TPCommunicator.PUSH_MAX_TOKENS`` tokens every pattern - quantisation and scale-factor layout included - is ONE
    kernel (one-shot push, :meth:`TPCommunicator.push_allreduce`).  Larger messages use the in-switch ``multimem.ld_reduce``
    kernel (AR + residual + RMSNorm + fp8) and, for NVFP4 only, the native quantiser as a second launch."""
    P = AllReduceFusionPattern
    comm = workspace.comm
    tokens, hidden = input.shape
    if pattern in (P.kMoEReductionARResidualRMSNorm, P.kMoEFinalizeARResidualRMSNorm):
        raise ValueError("use trtllm_moe_allreduce_fusion / trtllm_moe_finalize_allreduce_fusion for MoE patterns")
    if pattern != P.kAllReduce and (residual_in is None or rms_gamma is None):
        raise ValueError("residual_in and rms_gamma are required for the RMSNorm patterns")
    if input.is_cuda and use_oneshot is not False and comm.push_supported(tokens, hidden, input.dtype):
        if pattern == P.kAllReduce:
            out = output if output is not None else torch.empty_like(input)
            _push_fused(comm, input, pattern, allreduce_out=out, pdl=launch_with_pdl)
            return out
        res_out = residual_out if residual_out is not None else residual_in  # in place when no separate output is given
        need_norm = norm_out is not None or pattern in (P.kARResidualRMSNorm, P.kARResidualRMSNormOutFP8Quant, P.kARResidualRMSNormOutFP4Quant)
        if need_norm and norm_out is None:
            norm_out = torch.empty_like(input)
        if pattern in (P.kARResidualRMSNormFP8Quant, P.kARResidualRMSNormOutFP8Quant) and quant_out is None:
            quant_out = torch.empty(tokens, hidden, dtype=torch.float8_e4m3fn, device=input.device)
        if pattern in (P.kARResidualRMSNormFP4Quant, P.kARResidualRMSNormOutFP4Quant):
            if quant_out is None:
                quant_out = torch.empty(tokens, hidden // 2, dtype=torch.uint8, device=input.device)
            if scale_out is None:
                n_sf = compute_fp4_swizzled_layout_sf_size(tokens, hidden // 16) if layout_code != QuantizationSFLayout.LINEAR else tokens * hidden // 16
                scale_out = torch.empty(n_sf, dtype=torch.uint8, device=input.device)
        _push_fused(comm, input, pattern, allreduce_out=output, residual_in=residual_in, residual_out=res_out, rms_gamma=rms_gamma,
                    rms_eps=rms_eps, norm_out=norm_out, quant_out=quant_out, scale_out=scale_out, scale_factor=scale_factor,
                    layout_code=layout_code, pdl=launch_with_pdl)
        return norm_out if norm_out is not None else quant_out
    # ---- large messages: in-switch pull kernel
    if pattern == P.kAllReduce:
        return comm.allreduce_add_rmsnorm(input, None, None, out=output, enable_pdl=launch_with_pdl)
    res = residual_in if residual_out is None else residual_out
    if residual_out is not None and residual_out.data_ptr() != residual_in.data_ptr():
        residual_out.copy_(residual_in)
    norm = comm.allreduce_add_rmsnorm(input, res, rms_gamma, rms_eps, out=norm_out, two_shot=False, enable_pdl=launch_with_pdl)
    q = _quant_after(norm, pattern, quant_out, scale_out, scale_factor, layout_code)
    return q if q is not None and norm_out is None else norm


def trtllm_allreduce_fusion(allreduce_in: torch.Tensor, world_size: int, world_rank: int, token_num: int, hidden_dim: int,
                            workspace_ptrs, launch_with_pdl: bool, trigger_completion_at_end: bool, fp32_acc: bool,
                            pattern_code: int, use_oneshot: Optional[bool], allreduce_out: Optional[torch.Tensor],
                            residual_in: Optional[torch.Tensor], residual_out: Optional[torch.Tensor],
                            norm_out: Optional[torch.Tensor], quant_out: Optional[torch.Tensor],
                            scale_out: Optional[torch.Tensor], rms_gamma: Optional[torch.Tensor], rms_eps: Optional[float],
                            scale_factor=None, layout_code=None, metadata: Optional[dict] = None,
                            block_quant_group_size: Optional[int] = None) -> None:
    ws = workspace_ptrs[0] if isinstance(workspace_ptrs, (list, tuple)) else workspace_ptrs
    x = allreduce_in.view(token_num, hidden_dim)
    v = lambda t: t.view(token_num, hidden_dim) if t is not None else None  # noqa: E731
    if pattern_code in (AllReduceFusionPattern.kARResidualRMSNormPerTokenGroupFP8PackedQuant,
                        AllReduceFusionPattern.kARResidualRMSNormOutPerTokenGroupFP8PackedQuant):
        raise NotImplementedError("per-token-group fp8 packed quantisation pattern")
    big = not (x.is_cuda and use_oneshot is not False and ws.comm.push_supported(token_num, hidden_dim, x.dtype))
    allreduce_fusion(x, ws, pattern_code, launch_with_pdl, trigger_completion_at_end, v(allreduce_out), v(residual_out), v(norm_out),
                     quant_out, scale_out, v(residual_in), rms_gamma, rms_eps if rms_eps is not None else 1e-6, scale_factor, layout_code,
                     use_oneshot, fp32_acc)
    if big and allreduce_out is not None and residual_out is not None and residual_in is not None and pattern_code != AllReduceFusionPattern.kAllReduce:
        # pull kernel (large messages) has no separate raw-sum output: recover it from the residual stream
        allreduce_out.view(token_num, hidden_dim).copy_((residual_out.view(token_num, hidden_dim).float() -
                                                         residual_in.view(token_num, hidden_dim).float()).to(allreduce_out.dtype))


def trtllm_custom_all_reduce(inp: torch.Tensor, out: torch.Tensor, tp_size: int, tp_rank: int, token_num: int, fusion_op_code: int,
                             strategy_code, config_code, launch_with_pdl: bool, flag_value: int, peer_comm_buffer_ptrs,
                             peer_barrier_ptrs_in=None, peer_barrier_ptrs_out=None, bias: Optional[torch.Tensor] = None,
                             residual: Optional[torch.Tensor] = None, weight: Optional[torch.Tensor] = None,
                             weight_pre_residual_norm: Optional[torch.Tensor] = None, eps: Optional[float] = None,
                             intermediate_buffer: Optional[torch.Tensor] = None, lamport_peer_comm_buffer_ptrs_0=None,
                             lamport_peer_comm_buffer_ptrs_1=None, lamport_peer_comm_buffer_ptrs_2=None) -> None:
    """Legacy custom all-reduce in the reference's calling convention (trtllm_ar.py :809): ``out [token_num, hidden]`` receives the sum of
    ``inp`` over the TP group; with ``fusion_op_code = RESIDUAL_RMS_NORM`` it receives ``rmsnorm(sum + bias + residual) * weight`` and
    ``intermediate_buffer`` the pre-norm sum.  ``peer_comm_buffer_ptrs`` is a row (tensor or list) of the table returned by
    :func:`trtllm_create_ipc_workspace_for_all_reduce`, or that table / its workspace object.  The strategy / config codes pick between
    the reference's one-shot / two-shot kernels; the communicator chooses by message size here.  The other pointer rows and
    ``flag_value`` belong to the reference's barrier protocol (the communicator keeps its own epoch)."""
    ws = _resolve_legacy(peer_comm_buffer_ptrs)
    if ws.world_size != tp_size:
        raise ValueError(f"workspace was created for tp_size {ws.world_size}, called with {tp_size}")
    hidden = inp.numel() // token_num
    x = inp.view(token_num, hidden)
    o = out.view(token_num, hidden)
    op = int(fusion_op_code)
    if op == AllReduceFusionOp.NONE:
        ws.comm.allreduce_add_rmsnorm(x, None, None, out=o)
        return
    if op != AllReduceFusionOp.RESIDUAL_RMS_NORM:
        raise NotImplementedError(f"trtllm_custom_all_reduce: fusion_op_code {op} (use trtllm_allreduce_fusion for the quantising patterns)")
    if weight_pre_residual_norm is not None:
        raise NotImplementedError("trtllm_custom_all_reduce: weight_pre_residual_norm")
    if bias is not None:
        x = x + bias
    # the communicator updates its residual operand in place (residual += sum): that operand is the caller's intermediate_buffer
    # (the reference's output for the pre-norm sum) or a scratch copy - the caller's residual stays an input
    pre = intermediate_buffer.view(token_num, hidden) if intermediate_buffer is not None else torch.empty_like(x)
    pre.copy_(residual.view(token_num, hidden))
    ws.comm.allreduce_add_rmsnorm(x, pre, weight, eps or 1e-6, out=o, two_shot=False)


def _resolve_legacy(ptrs):
    if isinstance(ptrs, AllReduceFusionWorkspace):
        return ptrs
    if isinstance(ptrs, _IpcHandles):
        return ptrs.workspace
    first = ptrs
    while isinstance(first, (list, tuple)):
        first = first[0]
    if isinstance(first, AllReduceFusionWorkspace):
        return first
    key = int(first.flatten()[0]) if isinstance(first, torch.Tensor) else int(first)
    ws = _LEGACY.get(key)
    if ws is None:
        raise ValueError("peer_comm_buffer_ptrs does not come from trtllm_create_ipc_workspace_for_all_reduce (or the workspace was destroyed)")
    return ws


# ------------------------------------------------------------------ MoE fusions
def trtllm_moe_allreduce_fusion(world_size: int, world_rank: int, token_num: int, hidden_dim: int, workspace_ptrs,
                                launch_with_pdl: bool, residual_in: torch.Tensor, rms_gamma: torch.Tensor, rms_eps: float,
                                scale_factor, moe_reduction_device_num_experts: int, moe_reduction_scale_input: torch.Tensor,
                                moe_reduction_active_experts_token_input: torch.Tensor, moe_reduction_token_input: torch.Tensor,
                                layout_code=None, moe_allreduce_out: Optional[torch.Tensor] = None,
                                residual_out: Optional[torch.Tensor] = None, norm_out: Optional[torch.Tensor] = None,
                                quant_out: Optional[torch.Tensor] = None, scale_out: Optional[torch.Tensor] = None) -> None:
    """``x = sum_e scale[e, t] * active_experts_token[e, t, :] + token_input[t, :]`` reduced locally, then the fused
    all-reduce + residual + RMSNorm (reference trtllm_ar.py:1062)."""
    ws = workspace_ptrs[0] if isinstance(workspace_ptrs, (list, tuple)) else workspace_ptrs
    E = moe_reduction_device_num_experts
    if residual_in.is_cuda and ws.comm.push_supported(token_num, hidden_dim, residual_in.dtype):
        # ONE kernel: expert-weighted reduction -> one-shot push all-reduce -> + residual -> RMSNorm (-> NVFP4 quant)
        v = lambda t: t.view(token_num, hidden_dim) if t is not None else None  # noqa: E731
        _push_fused(ws.comm, None, AllReduceFusionPattern.kMoEReductionARResidualRMSNorm, allreduce_out=v(moe_allreduce_out),
                    residual_in=v(residual_in), residual_out=v(residual_out), rms_gamma=rms_gamma, rms_eps=rms_eps, norm_out=v(norm_out),
                    quant_out=quant_out, scale_out=scale_out, scale_factor=scale_factor, layout_code=layout_code, pdl=launch_with_pdl,
                    moe_reduction=(moe_reduction_active_experts_token_input.view(E, token_num, hidden_dim),
                                   moe_reduction_scale_input.view(E, token_num), moe_reduction_token_input.view(token_num, hidden_dim)))
        return
    act = moe_reduction_active_experts_token_input.view(E, token_num, hidden_dim).float()
    sc = moe_reduction_scale_input.view(E, token_num).float()
    x = (act * sc[..., None]).sum(0) + moe_reduction_token_input.view(token_num, hidden_dim).float()
    x = x.to(residual_in.dtype)
    res = residual_out if residual_out is not None else residual_in.clone()
    if residual_out is not None:
        residual_out.copy_(residual_in)
    norm = ws.comm.allreduce_add_rmsnorm(x, res.view(token_num, hidden_dim), rms_gamma, rms_eps,
                                         out=norm_out.view(token_num, hidden_dim) if norm_out is not None else None, two_shot=False)
    if moe_allreduce_out is not None:
        moe_allreduce_out.view(token_num, hidden_dim).copy_((res.view(token_num, hidden_dim).float() - residual_in.view(token_num, hidden_dim).float()).to(moe_allreduce_out.dtype))
    if quant_out is not None:
        _quant_after(norm, AllReduceFusionPattern.kARResidualRMSNormFP4Quant, quant_out, scale_out, scale_factor, layout_code)


def trtllm_moe_finalize_allreduce_fusion(allreduce_in: torch.Tensor, residual_in: torch.Tensor, norm_weight: torch.Tensor,
                                         expanded_idx_to_permuted_idx: torch.Tensor, norm_out: torch.Tensor,
                                         residual_out: torch.Tensor, quant_out: Optional[torch.Tensor], scale_out: Optional[torch.Tensor],
                                         workspace_ptrs, launch_with_pdl: bool, world_rank: int, world_size: int, eps: float,
                                         shared_expert_output: Optional[torch.Tensor] = None,
                                         expert_scale_factor: Optional[torch.Tensor] = None,
                                         routed_scaling_factor: Optional[float] = None, *, scale_factor=None, layout_code=None) -> None:
    """MoE finalize (top-k weighted un-permute, native kernel) + shared-expert add, then AR + residual + RMSNorm; argument order of the
    reference (trtllm_ar.py :1140).  ``routed_scaling_factor`` multiplies the routed sum before the shared expert is added - it is
    folded into the per-slot scales."""
    from .. import jit
    from ..utils import dtype_code, stream_ptr

    ws = workspace_ptrs[0] if isinstance(workspace_ptrs, (list, tuple)) else workspace_ptrs
    T, K = expanded_idx_to_permuted_idx.shape
    H = allreduce_in.shape[-1]
    if routed_scaling_factor is not None and float(routed_scaling_factor) != 1.0:
        base = expert_scale_factor.float() if expert_scale_factor is not None else torch.ones(T, K, device=allreduce_in.device)
        expert_scale_factor = base * float(routed_scaling_factor)
    if allreduce_in.is_cuda and ws.comm.push_supported(T, H, allreduce_in.dtype):
        # ONE kernel: top-k weighted un-permute (+ shared expert) -> one-shot push all-reduce -> + residual -> RMSNorm (-> NVFP4 quant)
        _push_fused(ws.comm, None, AllReduceFusionPattern.kMoEFinalizeARResidualRMSNorm, residual_in=residual_in.view(T, H),
                    residual_out=residual_out.view(T, H), rms_gamma=norm_weight, rms_eps=eps, norm_out=norm_out.view(T, H),
                    quant_out=quant_out, scale_out=scale_out, scale_factor=scale_factor, layout_code=layout_code, pdl=launch_with_pdl,
                    moe_finalize=(allreduce_in.reshape(-1, H), expanded_idx_to_permuted_idx, expert_scale_factor,
                                  shared_expert_output.view(T, H) if shared_expert_output is not None else None))
        return
    x = torch.empty(T, H, dtype=allreduce_in.dtype, device=allreduce_in.device) if shared_expert_output is None \
        else shared_expert_output.clone().view(T, H)
    w = expert_scale_factor.float().contiguous() if expert_scale_factor is not None else torch.ones(T, K, device=x.device)
    jit.load("moe").call("moe_finalize", allreduce_in.contiguous(), x, expanded_idx_to_permuted_idx.int().contiguous(), w, T, K, H,
                         0 if shared_expert_output is None else 1, dtype_code(x.dtype), 1, stream_ptr(x))
    residual_out.copy_(residual_in)
    norm = ws.comm.allreduce_add_rmsnorm(x, residual_out.view(T, H), norm_weight, eps, out=norm_out.view(T, H), two_shot=False)
    if quant_out is not None:
        _quant_after(norm, AllReduceFusionPattern.kARResidualRMSNormFP4Quant, quant_out, scale_out, scale_factor, layout_code)
