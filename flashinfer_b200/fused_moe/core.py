"""Fused Mixture-of-Experts.

Parity: reference flashinfer/fused_moe/core.py (trtllm_*_moe :2539-3503, cutlass_fused_moe :775-1009),
fused_routing_dsv3.py (fused_topk_deepseek), cute_dsl/fused_moe.py, tllm_enums.py.

Pipeline on B200 (all hand-written sm_100a kernels, PDL-chained):
  routing (9 methods, warp per token)  ->  sort / tile-padded permutation  ->  row gather  ->
  grouped tcgen05 GEMM (FC1, expert id = TMA coordinate)  ->  gated activation  ->  grouped GEMM (FC2)  ->
  finalize (top-k weighted un-permute).
Quantised entry points run natively: fp8 block-scale and fp8 per-tensor on the fp8 group-scaled tcgen05 grouped GEMM, NVFP4 on the
block-scaled (kind::mxf4nvf4) grouped GEMM; weight layout conversions (trtllm-gen shuffled / BlockMajorK) and the one format
without a tensor-core path (mxint4) are load-time preparations cached per weight tensor - nothing is de-quantised per call.  Gate/up convention follows the reference tests:
``h = x @ W1^T ; out = act(h[:, I:]) * h[:, :I]`` (second half is the gate).
"""
from __future__ import annotations

import weakref
from enum import IntEnum
from typing import List, Optional, Tuple, Union

import torch

from .. import jit
from ..activation import _act_and_mul
from ..utils import dtype_code, stream_ptr

_TILE = 128


class RoutingMethodType(IntEnum):
    Default = 0
    Renormalize = 1
    DeepSeekV3 = 2
    Llama4 = 3
    RenormalizeNaive = 4
    TopK = 5
    SigmoidRenorm = 6
    MiniMax2 = 7
    Sigmoid = 8
    Unspecified = 9


class ActivationType(IntEnum):
    Gelu = 0
    Relu = 1
    Silu = 2
    Swiglu = 3
    Geglu = 4
    SwigluBias = 5
    Relu2 = 6
    Identity = 7
    InvalidType = 8


class GatedActType(IntEnum):
    SwiGlu = 0
    GeGlu = 1


class WeightLayout(IntEnum):
    MajorK = 0
    MajorMn = 1
    BlockMajorK = 2


# ------------------------------------------------------------------ routing
class RoutingInputMode(IntEnum):
    """How routing information reaches a MoE entry point (reference fused_moe/core.py:81)."""
    FromLogits = 0
    PackedPrecomputed = 1
    UnpackedPrecomputed = 2


class Fp8QuantizationType(IntEnum):
    NoneFp8 = 0
    DeepSeekFp8 = 1
    MxFp8 = 2
    PerTensorFp8 = 3


class MoEInputs:
    """Flat container of the tensors a MoE runner consumes; field order = flat-list index (reference fused_moe/core.py:1011)."""
    _FIELDS = ("output", "routing_logits", "topk_ids", "expert_weights", "hidden_states", "hidden_states_scale",
               "per_token_scale")

    def __init__(self, output=None, routing_logits=None, topk_ids=None, expert_weights=None, hidden_states=None,
                 hidden_states_scale=None, per_token_scale=None):
        self.output, self.routing_logits, self.topk_ids, self.expert_weights = output, routing_logits, topk_ids, expert_weights
        self.hidden_states, self.hidden_states_scale, self.per_token_scale = hidden_states, hidden_states_scale, per_token_scale

    def to_list(self):
        return [getattr(self, n) for n in MoEInputs._FIELDS]

    @classmethod
    def from_list(cls, lst):
        return cls(**dict(zip(cls._FIELDS, lst)))

    @classmethod
    def idx(cls, name: str) -> int:
        return cls._FIELDS.index(name)


def route(routing_logits: torch.Tensor, routing_bias: Optional[torch.Tensor], top_k: int,
          routing_method_type: int = 0, n_group: Optional[int] = None, topk_group: Optional[int] = None,
          routed_scaling_factor: Optional[float] = None, norm_topk_prob: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns ``(topk_ids [T, K] int32, topk_weights [T, K] fp32)`` for any :class:`RoutingMethodType`."""
    T, E = routing_logits.shape
    method = int(routing_method_type)
    scale = float(routed_scaling_factor) if routed_scaling_factor is not None else 1.0
    if not routing_logits.is_cuda:
        return _route_cpu(routing_logits, routing_bias, top_k, method, n_group, topk_group, scale, norm_topk_prob)
    logits = routing_logits if routing_logits.dtype in (torch.float32, torch.bfloat16) else routing_logits.float()
    logits = logits.contiguous()
    ids = torch.empty(T, top_k, dtype=torch.int32, device=logits.device)
    w = torch.empty(T, top_k, dtype=torch.float32, device=logits.device)
    bias = routing_bias.float().contiguous() if routing_bias is not None else None
    jit.load("moe").call("moe_routing", logits, bias, ids, w, T, E, top_k, method, n_group or 1, topk_group or 1, scale,
                         1 if norm_topk_prob else 0, dtype_code(logits.dtype), 1, stream_ptr(logits))
    return ids, w


def _route_cpu(logits, bias, k, method, n_group, topk_group, scale, norm):
    x = logits.float()
    E = x.shape[1]
    if method in (0, 4):
        p = torch.softmax(x, -1)
        w, ids = torch.topk(p, k, -1)
        if method == 4:
            w = w / w.sum(-1, keepdim=True)
    elif method == 1:
        v, ids = torch.topk(x, k, -1)
        w = torch.softmax(v, -1)
    elif method == 3:
        v, ids = torch.topk(x, k, -1)
        w = torch.sigmoid(v)
    elif method == 5:
        w, ids = torch.topk(x, k, -1)
    elif method in (2, 6, 7, 8):
        s = torch.sigmoid(x)
        sel = s + (bias.float() if (bias is not None and method in (2, 7)) else 0)
        if method == 2 and n_group and n_group > 1:
            g = sel.view(-1, n_group, E // n_group)
            gs = g.topk(2, -1).values.sum(-1)
            keep = torch.zeros_like(gs, dtype=torch.bool).scatter_(1, gs.topk(topk_group, -1).indices, True)
            sel = torch.where(keep[..., None].expand_as(g).reshape(-1, E), sel, torch.full_like(sel, float("-inf")))
        ids = sel.topk(k, -1).indices
        w = s.gather(1, ids)
        if method in (2, 7):
            if norm or method == 7:
                w = w / (w.sum(-1, keepdim=True) + 1e-20)
            w = w * scale
        elif method == 6:
            w = w / (w.sum(-1, keepdim=True) + 1e-20)
    else:
        raise ValueError(f"unknown routing method {method}")
    return ids.int(), w.float()


def fused_topk_deepseek(scores: torch.Tensor, bias: torch.Tensor, n_group: int, topk_group: int, topk: int,
                        routed_scaling_factor: float, topk_values: Optional[torch.Tensor] = None,
                        topk_indices: Optional[torch.Tensor] = None, launch_with_pdl: bool = True,
                        routing_replay_out: Optional[torch.Tensor] = None):
    """DeepSeek-V3 no-aux-loss routing (reference fused_routing_dsv3.py): sigmoid + bias, grouped top-k.  ``routing_replay_out``
    (int16 ``[>= T, topk]``) also receives the selected expert ids."""
    ids, w = route(scores, bias, topk, RoutingMethodType.DeepSeekV3, n_group, topk_group, routed_scaling_factor, True)
    _replay(routing_replay_out, ids)
    if topk_values is not None:
        topk_values.copy_(w)
        w = topk_values
    if topk_indices is not None:
        topk_indices.copy_(ids)
        ids = topk_indices
    return w, ids


# ------------------------------------------------------------------ reference + pipeline
def moe_reference(x, topk_ids, topk_w, w1, w2, activation: str = "silu", local_expert_offset: int = 0):
    """fp32 oracle: ``w1 [E_local, 2I, H]``, ``w2 [E_local, H, I]``; gate = second half of FC1's output."""
    T, H = x.shape
    out = torch.zeros(T, H, dtype=torch.float32, device=x.device)
    E = w1.shape[0]
    inter = w2.shape[2]
    gated = w1.shape[1] == 2 * inter
    for e in range(E):
        sel = (topk_ids == e + local_expert_offset)
        tok, kk = torch.nonzero(sel, as_tuple=True)
        if tok.numel() == 0:
            continue
        h = x[tok].float() @ w1[e].float().t()
        if gated:
            up, gate = h[:, :inter], h[:, inter:]
            a = (torch.nn.functional.silu(gate) if activation == "silu" else torch.nn.functional.gelu(gate)) * up
        else:
            a = torch.nn.functional.silu(h) if activation == "silu" else torch.relu(h) ** 2
        y = a @ w2[e].float().t()
        out.index_add_(0, tok, y * topk_w[tok, kk].float()[:, None])
    return out


def moe_forward(x: torch.Tensor, topk_ids: torch.Tensor, topk_w: torch.Tensor, w1: torch.Tensor, w2: torch.Tensor,
                local_expert_offset: int = 0, num_experts: Optional[int] = None, activation: str = "silu",
                out: Optional[torch.Tensor] = None, do_finalize: bool = True):
    """Core bf16/fp16 MoE: x ``[T, H]``, w1 ``[E_local, 2I, H]``, w2 ``[E_local, H, I]``."""
    T, H = x.shape
    e_local, n1, _ = w1.shape
    inter = w2.shape[2]
    K = topk_ids.shape[1]
    if not x.is_cuda:
        res = moe_reference(x, topk_ids, topk_w, w1, w2, activation, local_expert_offset).to(x.dtype)
        if out is not None:
            out.copy_(res)
            return out
        return res
    if x.dtype not in (torch.float16, torch.bfloat16) or w1.dtype != x.dtype or w2.dtype != x.dtype:
        raise TypeError("moe_forward: x / w1 / w2 must share float16 or bfloat16")
    mod, gg = jit.load("moe"), jit.load("grouped_gemm_sm100")
    dev = x.device
    max_rows = (T * K + e_local * (_TILE - 1)) // _TILE * _TILE + _TILE
    max_tiles = max_rows // _TILE
    e2p = torch.empty(T * K, dtype=torch.int32, device=dev)
    p2t = torch.empty(max_rows, dtype=torch.int32, device=dev)
    tile_e = torch.empty(max_tiles, dtype=torch.int32, device=dev)
    offs = torch.empty(e_local + 1, dtype=torch.int32, device=dev)
    meta = torch.empty(4, dtype=torch.int32, device=dev)
    st = stream_ptr(x)
    ids = topk_ids.to(torch.int32).contiguous()
    ws = torch.empty(((T * K + 1023) // 1024) * e_local + 1, dtype=torch.int32, device=dev)
    mod.call("moe_sort", ids, T, K, num_experts or e_local, local_expert_offset, e_local, _TILE, max_rows, e2p, p2t,
             tile_e, offs, meta, ws, 1, st)
    xp = torch.empty(max_rows, H, dtype=x.dtype, device=dev)
    mod.call("moe_gather", x, xp, p2t, meta, max_rows, H, x.stride(0), 0, e2p, T * K, K, dtype_code(x.dtype), 1, st)
    h1 = torch.empty(max_rows, n1, dtype=x.dtype, device=dev)
    gg.call("grouped_gemm_nt", xp, w1.contiguous(), h1, tile_e, meta, max_tiles, n1, H, e_local, H, n1, p2t,
            dtype_code(x.dtype), 1, st)
    if n1 == 2 * inter:
        a = torch.empty(max_rows, inter, dtype=x.dtype, device=dev)
        _act_and_mul("silu" if activation == "silu" else "gelu", h1, a, True, gate_second=True, row_map=e2p, row_list=True)
    else:
        a = torch.nn.functional.silu(h1) if activation == "silu" else torch.relu(h1) ** 2
    h2 = torch.empty(max_rows, H, dtype=x.dtype, device=dev)
    gg.call("grouped_gemm_nt", a, w2.contiguous(), h2, tile_e, meta, max_tiles, H, inter, e_local, inter, H, p2t,
            dtype_code(x.dtype), 1, st)
    if not do_finalize:
        return h2, e2p, topk_w
    if out is None:
        out = torch.empty(T, H, dtype=x.dtype, device=dev)
    mod.call("moe_finalize", h2, out, e2p, topk_w.float().contiguous(), T, K, H, 0, dtype_code(x.dtype), 1, st)
    return out


_SF_CACHE: dict = {}
_UNIT: dict = {}


def _unit_scale(dev, v: float = 1.0) -> torch.Tensor:
    """Cached 1-element fp32 tensor (avoids a fill kernel per MoE call)."""
    key = (str(dev), float(v))
    t = _UNIT.get(key)
    if t is None:
        t = torch.full((1,), float(v), dtype=torch.float32, device=dev)
        _UNIT[key] = t
    return t


def _swizzle_expert_sf(sf: torch.Tensor, E: int, N: int, kc: int) -> torch.Tensor:
    """Per-expert linear ``[E, N, kc]`` UE4M3 scales -> 128x4-swizzled ``[E, bytes]`` (cached per weight tensor: static)."""
    from ..quantization.fp4 import _swizzled_sf_size, block_scale_interleave

    key = (sf.data_ptr(), E, N, kc)
    hit = _SF_CACHE.get(key)
    if hit is not None:
        return hit
    per = _swizzled_sf_size(N, kc)
    b = sf.view(torch.uint8)
    if b.numel() == E * per and not (b.dim() == 3 and b.shape[1:] == (N, kc) and per != N * kc):
        out = b.reshape(E, per) if b.dim() != 3 else block_scale_interleave(b.reshape(E, N, kc).contiguous()).reshape(E, per)
    else:
        out = block_scale_interleave(b.reshape(E, N, kc).contiguous()).reshape(E, per)
    if len(_SF_CACHE) > 64:
        _SF_CACHE.clear()
    _SF_CACHE[key] = out
    return out


def moe_forward_nvfp4(x: torch.Tensor, topk_ids: torch.Tensor, topk_w: torch.Tensor, w1_fp4: torch.Tensor, w1_sf: torch.Tensor,
                      w1_alpha, w2_fp4: torch.Tensor, w2_sf: torch.Tensor, w2_alpha, local_expert_offset: int = 0,
                      num_experts: Optional[int] = None, act_global_scale: float = 1.0, out: Optional[torch.Tensor] = None):
    """NVFP4 MoE on the block-scaled tcgen05 grouped GEMM (``kind::mxf4nvf4``): activations are quantised on the fly
    (per-16 UE4M3 scales in the 128x4 layout of the permuted matrix), ``w1_fp4 [E, 2I, H/2]`` / ``w2_fp4 [E, H, I/2]`` are packed
    e2m1 with linear ``[E, N, K/16]`` (or pre-swizzled) UE4M3 scales, ``w*_alpha [E]`` are the per-expert output de-quantisation
    scales (1 / (activation global scale * weight global scale))."""
    from ..gemm.lowp import grouped_gemm_nvfp4
    from ..quantization.fp4 import moe_fp4_quantize

    T, H = x.shape
    e_local, n1, _ = w1_fp4.shape
    inter = w2_fp4.shape[2] * 2
    K = topk_ids.shape[1]
    dev = x.device
    mod = jit.load("moe")
    max_rows = (T * K + e_local * (_TILE - 1)) // _TILE * _TILE + _TILE
    max_tiles = max_rows // _TILE
    e2p = torch.empty(T * K, dtype=torch.int32, device=dev)
    p2t = torch.empty(max_rows, dtype=torch.int32, device=dev)
    tile_e = torch.empty(max_tiles, dtype=torch.int32, device=dev)
    offs = torch.empty(e_local + 1, dtype=torch.int32, device=dev)
    meta = torch.empty(4, dtype=torch.int32, device=dev)
    st = stream_ptr(x)
    ids = topk_ids.to(torch.int32).contiguous()
    ws = torch.empty(((T * K + 1023) // 1024) * e_local + 1, dtype=torch.int32, device=dev)
    mod.call("moe_sort", ids, T, K, num_experts or e_local, local_expert_offset, e_local, _TILE, max_rows, e2p, p2t, tile_e, offs,
             meta, ws, 1, st)
    xb = x if x.dtype in (torch.float16, torch.bfloat16) else x.to(torch.bfloat16)
    if xb.stride(-1) != 1:
        xb = xb.contiguous()
    gs = _unit_scale(dev, act_global_scale)
    # gather + quantise in one kernel: token rows go straight into the permuted NVFP4 activation matrix
    xq, xsf = moe_fp4_quantize(xb, max_rows, H, p2t, gs, gather=True, gated=False, row_list=e2p, list_div=K)
    a1 = torch.as_tensor(w1_alpha, dtype=torch.float32, device=dev).reshape(-1)
    a2 = torch.as_tensor(w2_alpha, dtype=torch.float32, device=dev).reshape(-1)
    if a1.numel() == 1:
        a1 = a1.expand(e_local)
    if a2.numel() == 1:
        a2 = a2.expand(e_local)
    if act_global_scale != 1.0:
        a1, a2 = a1 / act_global_scale, a2 / act_global_scale
    a1, a2 = a1.contiguous(), a2.contiguous()
    sf1 = _swizzle_expert_sf(w1_sf, e_local, n1, H // 16)
    sf2 = _swizzle_expert_sf(w2_sf, e_local, H, inter // 16)
    h1 = grouped_gemm_nvfp4(xq, xsf, w1_fp4, sf1, a1, tile_e, meta, out_dtype=xb.dtype, row_map=p2t)
    # SwiGLU + quantisation of the FC2 input in one kernel
    aq, asf = moe_fp4_quantize(h1, max_rows, inter, p2t, gs, gather=False, gated=True, row_list=e2p)
    h2 = grouped_gemm_nvfp4(aq, asf, w2_fp4, sf2, a2, tile_e, meta, out_dtype=xb.dtype, row_map=p2t)
    if out is None:
        out = torch.empty(T, H, dtype=xb.dtype, device=dev)
    mod.call("moe_finalize", h2, out, e2p, topk_w.float().contiguous(), T, K, H, 0, dtype_code(xb.dtype), 1, st)
    return out


def moe_forward_fp8_block(x: torch.Tensor, x_scale: Optional[torch.Tensor], topk_ids: torch.Tensor, topk_w: torch.Tensor,
                          w1: torch.Tensor, w1_scale: torch.Tensor, w2: torch.Tensor, w2_scale: torch.Tensor,
                          local_expert_offset: int = 0, num_experts: Optional[int] = None,
                          out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """DeepSeek-V3 fp8 MoE on the native fp8 tensor-core path: ``w1 [E, 2I, H]`` / ``w2 [E, H, I]`` e4m3 with 128 x 128 fp32
    block scales ``[E, N/128, K/128]``; activations are quantised per 1 x 128 group on the fly.

    sort -> (gather + 1x128 quant) -> grouped fp8 GEMM (per-slab TMEM promotion) -> (SwiGLU + 1x128 quant) -> grouped fp8
    GEMM -> finalize: six kernels, no weight de-quantisation.  ``x`` may be bf16 / fp16, or e4m3 with ``x_scale [H/128, T]``
    (the trtllm layout).  Reference: trtllm_fp8_block_scale_moe (flashinfer/fused_moe/core.py)."""
    from ..gemm.lowp import fp8_group_quantize, grouped_gemm_fp8_groupwise

    T, H = x.shape
    e_local, n1, _ = w1.shape
    inter = w2.shape[2]
    K = topk_ids.shape[1]
    if x.dtype == torch.float8_e4m3fn:
        s = x_scale.float().t().repeat_interleave(128, -1)[:, :H]
        x = (x.float() * s).to(torch.bfloat16)
    if (not x.is_cuda) or H % 128 or inter % 128 or n1 != 2 * inter:
        res = moe_forward(x, topk_ids, topk_w, _dequant_fp8_block(w1, w1_scale, dtype=x.dtype),
                          _dequant_fp8_block(w2, w2_scale, dtype=x.dtype), local_expert_offset, num_experts)
        if out is not None:
            out.copy_(res)
            return out
        return res
    dev = x.device
    mod = jit.load("moe")
    max_rows = (T * K + e_local * (_TILE - 1)) // _TILE * _TILE + _TILE
    e2p = torch.empty(T * K, dtype=torch.int32, device=dev)
    p2t = torch.empty(max_rows, dtype=torch.int32, device=dev)
    tile_e = torch.empty(max_rows // _TILE, dtype=torch.int32, device=dev)
    offs = torch.empty(e_local + 1, dtype=torch.int32, device=dev)
    meta = torch.empty(4, dtype=torch.int32, device=dev)
    st = stream_ptr(x)
    ids = topk_ids.to(torch.int32).contiguous()
    ws = torch.empty(((T * K + 1023) // 1024) * e_local + 1, dtype=torch.int32, device=dev)
    mod.call("moe_sort", ids, T, K, num_experts or e_local, local_expert_offset, e_local, _TILE, max_rows, e2p, p2t, tile_e, offs,
             meta, ws, 1, st)
    xq, xs = fp8_group_quantize(x, max_rows, gated=False, row_list=e2p, gather=True, list_div=K)
    h1 = grouped_gemm_fp8_groupwise(xq, xs, w1, w1_scale, tile_e, meta, x.dtype, row_map=p2t)
    aq, a_s = fp8_group_quantize(h1, max_rows, gated=True, row_list=e2p)
    h2 = grouped_gemm_fp8_groupwise(aq, a_s, w2, w2_scale, tile_e, meta, x.dtype, row_map=p2t)
    if out is None:
        out = torch.empty(T, H, dtype=x.dtype, device=dev)
    mod.call("moe_finalize", h2, out, e2p, topk_w.float().contiguous(), T, K, H, 0, dtype_code(x.dtype), 1, st)
    return out


def reorder_rows_for_gated_act_gemm(x: torch.Tensor) -> torch.Tensor:
    """Interleave the two halves of the rows ([up | gate] -> u0 g0 u1 g1 ...), the weight layout that lets a
    GEMM epilogue see matching up/gate columns in one tile (reference core.py:133)."""
    m = x.shape[0]
    half = m // 2
    return torch.stack([x[:half], x[half:]], 1).reshape(x.shape)


# ------------------------------------------------------------------ de-quantisation helpers
def _dequant_fp8_block(w: torch.Tensor, scale: torch.Tensor, block: int = 128, dtype=torch.bfloat16) -> torch.Tensor:
    """w [..., N, K] fp8, scale [..., N/block, K/block] fp32 (DeepSeek 128x128 weight blocks)."""
    s = scale.float().repeat_interleave(block, -2).repeat_interleave(block, -1)[..., : w.shape[-2], : w.shape[-1]]
    return (w.float() * s).to(dtype)


def _dequant_nvfp4(w: torch.Tensor, sf: torch.Tensor, global_scale, vec: int = 16, dtype=torch.bfloat16) -> torch.Tensor:
    """w [E, N, K/2] uint8 (e2m1 pairs), sf [E, N, K/vec] UE4M3 bytes (linear layout), global scale per expert."""
    from ..quantization.fp4 import E2M1_VALUES

    lut = torch.tensor(E2M1_VALUES + [-v for v in E2M1_VALUES], device=w.device)
    wb = w.view(torch.uint8)
    vals = torch.stack([lut[(wb & 0xF).long()], lut[(wb >> 4).long()]], -1).flatten(-2)
    s = sf.view(torch.uint8).view(torch.float8_e4m3fn).float().reshape(*vals.shape[:-1], -1)
    s = s.repeat_interleave(vec, -1)[..., : vals.shape[-1]]
    g = torch.as_tensor(global_scale, device=w.device, dtype=torch.float32).reshape(-1, *([1] * (vals.ndim - 1)))
    return (vals * s * g).to(dtype)


# ------------------------------------------------------------------ load-time weight preparation (cached per weight tensor)
_PREP_CACHE: dict = {}


def _prepared(tag: str, tensors, fn):
    """One-time transformation of STATIC expert weights, cached on (tag, data_ptr, shape, dtype, version) of the inputs: layout
    conversions (un-shuffle, BlockMajorK -> MajorK) and the few formats without a native tensor-core path (mxint4) are prepared
    on first use - never per call (the quantised entry points used to de-quantise every expert weight in eager
    torch on every call).

    An entry is only trusted while the STORAGE it was computed from is still alive (weak references to the storage objects, which
    PyTorch keeps for as long as any tensor - parameter, ``.data`` alias, view - uses the memory): once it is freed the allocator may
    hand the same address to different weights of the same shape, and the key alone would match them."""
    live = [t for t in tensors if isinstance(t, torch.Tensor)]
    key = (tag,) + tuple((t.data_ptr(), tuple(t.shape), tuple(t.stride()), str(t.dtype), t._version) for t in live)
    hit = _PREP_CACHE.get(key)
    if hit is not None and all(r() is not None for r in hit[0]):
        return hit[1]
    if len(_PREP_CACHE) > 256:
        _PREP_CACHE.clear()
    value = fn()
    _PREP_CACHE[key] = ([weakref.ref(t.untyped_storage()) for t in live], value)
    return value


def _scale_tag(tag: str, scale) -> str:
    """Cache tag that carries a python-number global scale (tensor scales are part of the key through the tensor list)."""
    return tag if (scale is None or isinstance(scale, torch.Tensor)) else f"{tag}:{float(scale)!r}"


def _shuffle_block_rows(m: int, epilogue_tile_m: int) -> torch.Tensor:
    """``row_indices[new_row] = old_row`` of the trtllm-gen ``shuffle_matrix_a`` row permutation: rows move inside blocks of 16
    (32 when ``epilogue_tile_m % 128 == 0``) so that row ``i`` of a block lands at ``(i % (B / 8)) * 8 + i // (B / 8)``."""
    b = 32 if epilogue_tile_m % 128 == 0 else 16
    if m % b:
        raise ValueError(f"shuffled weights need a row count that is a multiple of {b}")
    old = torch.arange(m)
    i = old % b
    new = (old // b) * b + (i % (b // 8)) * 8 + i // (b // 8)
    idx = torch.empty(m, dtype=torch.long)
    idx[new] = old
    return idx


def _unshuffle_expert_weights(w: torch.Tensor, epilogue_tile_m: int, block_major_k: bool, gated_interleaved: bool,
                              block_k_bytes: int = 128) -> torch.Tensor:
    """Undo the trtllm-gen weight pre-processing (reference fused_moe/core.py:133-233, tests/moe/test_trtllm_gen_fused_moe.py:1004):
    ``[E, K/bk, N, bk]`` BlockMajorK -> ``[E, N, K]``, inverse ``shuffle_matrix_a`` row permutation, inverse gated-row interleave
    (``reorder_rows_for_gated_act_gemm``).  The TMA-fed grouped GEMM here needs none of them."""
    raw = w.view(torch.uint8)
    if block_major_k:
        e, kb, n, bk = raw.shape
        raw = raw.permute(0, 2, 1, 3).reshape(e, n, kb * bk)
    e, n, kbytes = raw.shape
    idx = _shuffle_block_rows(n, epilogue_tile_m).to(raw.device)      # new -> old
    out = torch.empty_like(raw)
    out[:, idx] = raw                                                   # old row <- new row
    if gated_interleaved:                                               # rows (r0, rN/2, r1, ...) -> [first half | second half]
        out = torch.cat([out[:, 0::2], out[:, 1::2]], 1)
    return out.contiguous().view(w.dtype) if w.dtype != torch.uint8 else out.contiguous()


def _plain_weights(name: str, w: torch.Tensor, use_shuffled_weight: bool, weight_layout: int, epilogue_tile_m: int,
                   gated_interleaved: bool = False) -> torch.Tensor:
    bmk = int(weight_layout) == int(WeightLayout.BlockMajorK)
    if int(weight_layout) == int(WeightLayout.MajorMn):
        raise NotImplementedError(f"{name}: WeightLayout.MajorMn is not supported")
    if not use_shuffled_weight and not bmk:
        return w
    if bmk and not use_shuffled_weight:
        return _prepared("bmk", [w], lambda: _unshuffle_expert_weights_noperm(w))
    return _prepared(f"unshuffle{epilogue_tile_m}{int(bmk)}{int(gated_interleaved)}", [w],
                     lambda: _unshuffle_expert_weights(w, epilogue_tile_m, bmk, gated_interleaved))


def _unshuffle_expert_weights_noperm(w: torch.Tensor) -> torch.Tensor:
    raw = w.view(torch.uint8)
    e, kb, n, bk = raw.shape
    out = raw.permute(0, 2, 1, 3).reshape(e, n, kb * bk).contiguous()
    return out.view(w.dtype) if w.dtype != torch.uint8 else out


def _scalar_block_scales(scale: torch.Tensor, e: int, n: int, k: int) -> torch.Tensor:
    """Per-expert scalar de-quantisation scale as a (stride-0) ``[E, N/128, K/128]`` block-scale view: lets fp8 per-tensor weights
    run on the native fp8 group-scaled tensor-core pipeline without touching the weights."""
    return scale.float().reshape(-1, 1, 1).expand(e, (n + 127) // 128, (k + 127) // 128)


def moe_forward_fp8_per_tensor(x: torch.Tensor, topk_ids: torch.Tensor, topk_w: torch.Tensor, w1: torch.Tensor, alpha1,
                               w2: torch.Tensor, alpha2, local_expert_offset: int = 0, num_experts: Optional[int] = None,
                               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp8 (e4m3) expert weights with ONE de-quantisation scale per expert and GEMM: ``h = (x @ W1^T) * alpha1[e]``,
    ``out = (swiglu(h) @ W2^T) * alpha2[e]``.  Runs on the fp8 tensor-core MoE pipeline (:func:`moe_forward_fp8_block`): the
    scalar scales become stride-0 block-scale views and the activations are quantised per 1 x 128 group on the fly (finer than
    the per-tensor activation scale of the reference kernels, so at least as accurate).  ``x`` may already be e4m3 (its values
    are used as they are: the caller's activation scale is part of ``alpha1``, reference contract)."""
    e_local, n1, h = w1.shape
    inter = w2.shape[2]
    dev = x.device
    a1 = torch.as_tensor(alpha1, dtype=torch.float32, device=dev).reshape(-1)
    a2 = torch.as_tensor(alpha2, dtype=torch.float32, device=dev).reshape(-1)
    if a1.numel() == 1:
        a1 = a1.expand(e_local)
    if a2.numel() == 1:
        a2 = a2.expand(e_local)
    xb = x.to(torch.bfloat16) if x.dtype not in (torch.float16, torch.bfloat16) else x
    if (not x.is_cuda) or h % 128 or inter % 128 or n1 != 2 * inter:
        w1d = _prepared("fp8pt", [w1, a1], lambda: (w1.float() * a1.reshape(-1, 1, 1)).to(xb.dtype))
        w2d = _prepared("fp8pt", [w2, a2], lambda: (w2.float() * a2.reshape(-1, 1, 1)).to(xb.dtype))
        return moe_forward(xb, topk_ids, topk_w, w1d, w2d, local_expert_offset, num_experts, out=out)
    return moe_forward_fp8_block(xb, None, topk_ids, topk_w, w1, _scalar_block_scales(a1, e_local, n1, h), w2,
                                 _scalar_block_scales(a2, e_local, h, inter), local_expert_offset, num_experts, out=out)


# ------------------------------------------------------------------ trtllm-gen style entry points

def _activation_name(activation_type) -> str:
    a = int(activation_type)
    if a in (int(ActivationType.Swiglu), int(ActivationType.Silu)):
        return "silu"
    if a in (int(ActivationType.Geglu), int(ActivationType.Gelu)):
        return "gelu"
    if a == int(ActivationType.Relu2):
        return "relu2"
    raise NotImplementedError(f"activation_type {ActivationType(a).name} is not implemented")


def _fold_gate_scale(s1, s_gate, s2):
    """Per-expert output scales of the trtllm-gen MoE contract -> (alpha1, alpha2) of a pipeline that applies ONE scale to
    both halves of FC1.  Contract (reference tests/moe/test_trtllm_gen_fused_moe.py): ``act = silu(gate * s_gate) * (up * s1)``,
    ``out = (act @ W2^T) * s2``.  The up factor is linear, so ``s1 / s_gate`` moves to FC2's output scale exactly:
    alpha1 = s_gate on both halves, alpha2 = s2 * s1 / s_gate.  Device-side [E] arithmetic, no host sync."""
    if s_gate is None:
        return s1, s2
    dev = next((t.device for t in (s_gate, s1, s2) if isinstance(t, torch.Tensor)), None)
    s1t = torch.as_tensor(s1, dtype=torch.float32, device=dev)
    sg = torch.as_tensor(s_gate, dtype=torch.float32, device=dev)
    s2t = torch.as_tensor(s2, dtype=torch.float32, device=dev)
    return sg, s2t * (s1t / sg)


def _reject_unsupported(name: str, **kw):
    """Arguments the native pipelines do not implement must fail loudly instead of being ignored."""
    bad = [k for k, v in kw.items() if v is not None and v is not False]
    if bad:
        raise NotImplementedError(f"{name}: unsupported argument(s) {bad}")

def trtllm_bf16_moe(routing_logits, routing_bias, hidden_states, gemm1_weights, gemm2_weights, num_experts, top_k,
                    n_group, topk_group, intermediate_size, local_expert_offset, local_num_experts,
                    routed_scaling_factor=None, routing_method_type: int = 0, use_shuffled_weight: bool = True,
                    weight_layout: int = WeightLayout.BlockMajorK, do_finalize: bool = True, enable_pdl: bool = True,
                    tune_max_num_tokens: int = 8192, activation_type: int = ActivationType.Swiglu.value,
                    norm_topk_prob: bool = True, routing_replay_out=None):
    """bf16 MoE with fused routing.  The defaults are the reference's (core.py :2560): weights pre-processed for trtllm-gen
    (``use_shuffled_weight=True``, ``WeightLayout.BlockMajorK``) - they are converted back ONCE per tensor (cached).  The TMA-fed
    grouped GEMM itself needs no pre-shuffling: pass ``use_shuffled_weight=False, weight_layout=WeightLayout.MajorK`` with plain
    K-major ``[E_local, 2I, H]`` / ``[E_local, H, I]`` weights to skip the conversion."""
    # trtllm-gen pre-processed weights (gated-row interleave + shuffle_matrix_a(epilogue_tile_m=128) [+ BlockMajorK]) are
    # converted back to plain MajorK ONCE per weight tensor (cached); the TMA-fed grouped GEMM needs no pre-shuffle
    gemm1_weights = _plain_weights("trtllm_bf16_moe", gemm1_weights, use_shuffled_weight, weight_layout, 128, gated_interleaved=True)
    gemm2_weights = _plain_weights("trtllm_bf16_moe", gemm2_weights, use_shuffled_weight, weight_layout, 128)
    ids, w = route(routing_logits, routing_bias, top_k, routing_method_type, n_group, topk_group, routed_scaling_factor,
                   norm_topk_prob)
    _replay(routing_replay_out, ids)
    return moe_forward(hidden_states, ids, w, gemm1_weights, gemm2_weights, local_expert_offset, num_experts,
                       activation=_activation_name(activation_type), do_finalize=do_finalize)


def _replay(routing_replay_out: Optional[torch.Tensor], ids: torch.Tensor) -> None:
    """Routing replay (reference fused_routing_dsv3.py :45): the selected expert ids of every token, int16 ``[>= T, top_k]``."""
    if routing_replay_out is None:
        return
    t, k = ids.shape
    if routing_replay_out.dtype != torch.int16:
        raise ValueError(f"routing_replay_out must be int16, got {routing_replay_out.dtype}")
    if routing_replay_out.shape[0] < t or routing_replay_out.shape[1] != k:
        raise ValueError(f"routing_replay_out shape[0] must be >= {t} and shape[1] must be {k}, got {tuple(routing_replay_out.shape)}")
    routing_replay_out[:t].copy_(ids)


_TAIL_KEYS = frozenset({"do_finalize", "enable_pdl", "tune_max_num_tokens", "activation_type", "norm_topk_prob", "routing_replay_out", "output",
                        "per_token_scale", "fp8_quantization_type", "use_shuffled_weight", "weight_layout"})


def _tail(name: str, kw: dict, ids: Optional[torch.Tensor] = None, supports: Tuple[str, ...] = ()) -> None:
    """Keyword tail of the trtllm_* entry points (the reference's trailing optional arguments).  Launch / tuning hints (``enable_pdl``,
    ``tune_max_num_tokens``) do not change the result and are accepted; ``routing_replay_out`` is written; anything that would
    change the result and is not implemented by this entry point is refused, and an unknown name is a TypeError as usual."""
    unknown = set(kw) - _TAIL_KEYS
    if unknown:
        raise TypeError(f"{name}() got unexpected keyword argument(s) {sorted(unknown)}")
    bad = []
    if kw.get("do_finalize", True) is False and "do_finalize" not in supports:
        bad.append("do_finalize=False")
    act = kw.get("activation_type")
    if act is not None and int(act) != int(ActivationType.Swiglu) and "activation_type" not in supports:
        bad.append(f"activation_type={ActivationType(int(act)).name}")
    if kw.get("per_token_scale") is not None:
        bad.append("per_token_scale")
    fq = kw.get("fp8_quantization_type")
    if fq is not None and int(fq) != int(Fp8QuantizationType.DeepSeekFp8):
        bad.append(f"fp8_quantization_type={Fp8QuantizationType(int(fq)).name}")
    if (kw.get("use_shuffled_weight") or kw.get("weight_layout")) and "weights" not in supports:
        bad.append("use_shuffled_weight / weight_layout")
    if bad:
        raise NotImplementedError(f"{name}: unsupported argument(s) {bad}")
    if ids is not None:
        _replay(kw.get("routing_replay_out"), ids)


def _into(output: Optional[torch.Tensor], res):
    """``output=`` of the trtllm_* entry points: the finalized result is written into the caller's buffer."""
    if output is None:
        return res
    output.copy_(res[0] if isinstance(res, (list, tuple)) else res)
    return output


def _unpack_routed(topk_ids: torch.Tensor):
    """Packed routing used by the *_routed_moe APIs: int32 = (expert_id << 16) | bf16(weight)."""
    ids = (topk_ids >> 16).int()
    w = (topk_ids & 0xFFFF).to(torch.int16).view(torch.bfloat16).float()
    return ids, w


def trtllm_bf16_routed_moe(topk_ids, hidden_states, gemm1_weights, gemm2_weights, num_experts, top_k, n_group,
                           topk_group, intermediate_size, local_expert_offset, local_num_experts,
                           routed_scaling_factor=None, routing_method_type: int = 0, use_shuffled_weight: bool = True,
                           weight_layout: int = WeightLayout.BlockMajorK, do_finalize: bool = True, enable_pdl: bool = True,
                           tune_max_num_tokens: int = 8192, activation_type: int = ActivationType.Swiglu.value,
                           routing_replay_out=None):
    """Pre-routed bf16 MoE (packed ``(expert << 16) | bf16 weight`` words); weight defaults as in :func:`trtllm_bf16_moe`.
    ``routing_replay_out [>= T, top_k]`` int16 receives the expert ids."""
    ids, w = _unpack_routed(topk_ids)
    _replay(routing_replay_out, ids)
    gemm1_weights = _plain_weights("trtllm_bf16_routed_moe", gemm1_weights, use_shuffled_weight, weight_layout, 128, gated_interleaved=True)
    gemm2_weights = _plain_weights("trtllm_bf16_routed_moe", gemm2_weights, use_shuffled_weight, weight_layout, 128)
    return moe_forward(hidden_states, ids, w, gemm1_weights, gemm2_weights, local_expert_offset, num_experts,
                       activation=_activation_name(activation_type), do_finalize=do_finalize)


def trtllm_fp8_per_tensor_scale_moe(routing_logits, routing_bias, hidden_states, gemm1_weights, output1_scales_scalar,
                                    output1_scales_gate_scalar, gemm2_weights, output2_scales_scalar, num_experts, top_k,
                                    n_group, topk_group, intermediate_size, local_expert_offset, local_num_experts,
                                    routed_scaling_factor, use_routing_scales_on_input: bool = False,
                                    routing_method_type: int = 0, **kw):
    """fp8 per-tensor MoE on the native fp8 tensor-core pipeline (:func:`moe_forward_fp8_per_tensor`); no weight de-quantisation."""
    _reject_unsupported("trtllm_fp8_per_tensor_scale_moe", use_routing_scales_on_input=use_routing_scales_on_input)
    ids, w = route(routing_logits, routing_bias, top_k, routing_method_type, n_group, topk_group, routed_scaling_factor,
                   kw.get("norm_topk_prob", True))
    _tail("trtllm_fp8_per_tensor_scale_moe", kw, ids)
    a1, a2 = _fold_gate_scale(output1_scales_scalar, output1_scales_gate_scalar, output2_scales_scalar)
    return moe_forward_fp8_per_tensor(hidden_states, ids, w, gemm1_weights, a1, gemm2_weights, a2, local_expert_offset, num_experts)


def trtllm_fp8_block_scale_moe(routing_logits, routing_bias, hidden_states, hidden_states_scale, gemm1_weights,
                               gemm1_weights_scale, gemm2_weights, gemm2_weights_scale, num_experts, top_k, n_group,
                               topk_group, intermediate_size, local_expert_offset, local_num_experts,
                               routed_scaling_factor, routing_method_type: int = 0, use_shuffled_weight: bool = False,
                               weight_layout: int = 0, **kw):
    """DeepSeek-style fp8 (1x128 activation scales ``[H/128, T]``, 128x128 weight scales).  ``use_shuffled_weight`` / BlockMajorK weights
    (shuffle_matrix_a with epilogue_tile_m = 64, reference tests/moe/test_dpsk_fused_moe_fp8.py:684) are converted back once."""
    gemm1_weights = _plain_weights("trtllm_fp8_block_scale_moe", gemm1_weights, use_shuffled_weight, weight_layout, 64)
    gemm2_weights = _plain_weights("trtllm_fp8_block_scale_moe", gemm2_weights, use_shuffled_weight, weight_layout, 64)
    ids, w = route(routing_logits, routing_bias, top_k, routing_method_type, n_group, topk_group, routed_scaling_factor,
                   kw.get("norm_topk_prob", True))
    _tail("trtllm_fp8_block_scale_moe", kw, ids)
    return moe_forward_fp8_block(hidden_states, hidden_states_scale, ids, w, gemm1_weights, gemm1_weights_scale, gemm2_weights,
                                 gemm2_weights_scale, local_expert_offset, num_experts)


def trtllm_fp8_block_scale_routed_moe(topk_ids, routing_bias, hidden_states, hidden_states_scale, gemm1_weights,
                                      gemm1_weights_scale, gemm2_weights, gemm2_weights_scale, num_experts, top_k,
                                      n_group, topk_group, intermediate_size, local_expert_offset, local_num_experts,
                                      routed_scaling_factor, routing_method_type: int = 0, use_shuffled_weight: bool = False,
                                      weight_layout: int = 0, **kw):
    """Pre-routed form of :func:`trtllm_fp8_block_scale_moe` (packed ``(expert << 16) | bf16 weight`` words)."""
    gemm1_weights = _plain_weights("trtllm_fp8_block_scale_routed_moe", gemm1_weights, use_shuffled_weight, weight_layout, 64)
    gemm2_weights = _plain_weights("trtllm_fp8_block_scale_routed_moe", gemm2_weights, use_shuffled_weight, weight_layout, 64)
    ids, w = _unpack_routed(topk_ids)
    _tail("trtllm_fp8_block_scale_routed_moe", kw, ids)
    return _into(kw.get("output"), moe_forward_fp8_block(hidden_states, hidden_states_scale, ids, w, gemm1_weights, gemm1_weights_scale,
                                                         gemm2_weights, gemm2_weights_scale, local_expert_offset, num_experts))


def trtllm_fp4_block_scale_moe(routing_logits, routing_bias, hidden_states, hidden_states_scale, gemm1_weights,
                               gemm1_weights_scale, gemm1_bias, gemm1_alpha, gemm1_beta, gemm1_clamp_limit, gemm2_weights,
                               gemm2_weights_scale, gemm2_bias, output1_scale_scalar, output1_scale_gate_scalar,
                               output2_scale_scalar, num_experts, top_k, n_group, topk_group, intermediate_size,
                               local_expert_offset, local_num_experts, routed_scaling_factor,
                               routing_method_type: int = 0, do_finalize: bool = True, **kw):
    """NVFP4 weights (``[E, N, K/2]`` packed e2m1 + linear UE4M3 block scales) with bf16 or nvfp4 activations."""
    _reject_unsupported("trtllm_fp4_block_scale_moe", gemm1_bias=gemm1_bias, gemm1_alpha=gemm1_alpha, gemm1_beta=gemm1_beta,
                        gemm1_clamp_limit=gemm1_clamp_limit, gemm2_bias=gemm2_bias)
    ids, w = route(routing_logits, routing_bias, top_k, routing_method_type, n_group, topk_group, routed_scaling_factor,
                   kw.get("norm_topk_prob", True))
    _tail("trtllm_fp4_block_scale_moe", kw, ids, supports=("do_finalize",))
    return _into(kw.get("output"), _fp4_block_scale_core(hidden_states, hidden_states_scale, ids, w, gemm1_weights, gemm1_weights_scale, gemm2_weights,
                                                         gemm2_weights_scale, output1_scale_scalar, output1_scale_gate_scalar, output2_scale_scalar,
                                                         num_experts, intermediate_size, local_expert_offset, do_finalize))


def _fp4_block_scale_core(hidden_states, hidden_states_scale, ids, w, gemm1_weights, gemm1_weights_scale, gemm2_weights, gemm2_weights_scale,
                          output1_scale_scalar, output1_scale_gate_scalar, output2_scale_scalar, num_experts, intermediate_size,
                          local_expert_offset, do_finalize):
    """Shared body of the fused-routing and pre-routed NVFP4 entry points (routing already resolved to ``ids`` / ``w``)."""
    x = hidden_states
    if x.dtype == torch.uint8:
        x = _dequant_nvfp4(x, hidden_states_scale, 1.0)
    g1 = output1_scale_scalar if output1_scale_scalar is not None else 1.0
    g2 = output2_scale_scalar if output2_scale_scalar is not None else 1.0
    g1, g2 = _fold_gate_scale(g1, output1_scale_gate_scalar, g2)
    if x.is_cuda and do_finalize and hidden_states.shape[-1] % 64 == 0 and intermediate_size % 64 == 0:
        return moe_forward_nvfp4(x, ids, w, gemm1_weights, gemm1_weights_scale, g1, gemm2_weights, gemm2_weights_scale, g2,
                                 local_expert_offset, num_experts)
    w1 = _prepared(_scale_tag("nvfp4deq", g1), [gemm1_weights, gemm1_weights_scale, g1], lambda: _dequant_nvfp4(gemm1_weights, gemm1_weights_scale, g1))
    w2 = _prepared(_scale_tag("nvfp4deq", g2), [gemm2_weights, gemm2_weights_scale, g2], lambda: _dequant_nvfp4(gemm2_weights, gemm2_weights_scale, g2))
    return moe_forward(x.to(torch.bfloat16), ids, w, w1, w2, local_expert_offset, num_experts, do_finalize=do_finalize)


def trtllm_fp4_block_scale_routed_moe(topk_ids, routing_bias, hidden_states, hidden_states_scale, gemm1_weights,
                                      gemm1_weights_scale, gemm1_bias, gemm1_alpha, gemm1_beta, gemm1_clamp_limit,
                                      gemm2_weights, gemm2_weights_scale, gemm2_bias, output1_scale_scalar,
                                      output1_scale_gate_scalar, output2_scale_scalar, num_experts, top_k, n_group,
                                      topk_group, intermediate_size, local_expert_offset, local_num_experts,
                                      routed_scaling_factor, routing_method_type: int = 0, do_finalize: bool = True, **kw):
    """Pre-routed form of :func:`trtllm_fp4_block_scale_moe` (packed ``(expert << 16) | bf16 weight`` words)."""
    _reject_unsupported("trtllm_fp4_block_scale_routed_moe", gemm1_bias=gemm1_bias, gemm1_alpha=gemm1_alpha,
                        gemm1_beta=gemm1_beta, gemm1_clamp_limit=gemm1_clamp_limit, gemm2_bias=gemm2_bias)
    ids, w = _unpack_routed(topk_ids)
    _tail("trtllm_fp4_block_scale_routed_moe", kw, ids, supports=("do_finalize",))
    return _into(kw.get("output"), _fp4_block_scale_core(hidden_states, hidden_states_scale, ids, w, gemm1_weights, gemm1_weights_scale, gemm2_weights,
                                                         gemm2_weights_scale, output1_scale_scalar, output1_scale_gate_scalar, output2_scale_scalar,
                                                         num_experts, intermediate_size, local_expert_offset, do_finalize))


def trtllm_mxint4_block_scale_moe(routing_logits, routing_bias, hidden_states, gemm1_weights, gemm1_weights_scale,
                                  gemm1_alpha, gemm1_beta, gemm1_clamp_limit, gemm2_weights, gemm2_weights_scale,
                                  num_experts, top_k, n_group, topk_group, intermediate_size, local_expert_offset,
                                  local_num_experts, routed_scaling_factor, routing_method_type: int = 0, **kw):
    """MXINT4 weights: int4 pairs in uint8 + bf16 scales per 32 elements."""
    ids, w = route(routing_logits, routing_bias, top_k, routing_method_type, n_group, topk_group, routed_scaling_factor,
                   kw.get("norm_topk_prob", True))
    _tail("trtllm_mxint4_block_scale_moe", kw, ids)

    def deq(wq, sc):
        b = wq.view(torch.uint8)
        lo = (b & 0xF).to(torch.int8)
        hi = (b >> 4).to(torch.int8)
        lo = torch.where(lo > 7, lo - 16, lo)
        hi = torch.where(hi > 7, hi - 16, hi)
        vals = torch.stack([lo, hi], -1).flatten(-2).float()
        s = sc.float().repeat_interleave(32, -1)[..., : vals.shape[-1]]
        return (vals * s).to(torch.bfloat16)

    _reject_unsupported("trtllm_mxint4_block_scale_moe", gemm1_alpha=gemm1_alpha, gemm1_beta=gemm1_beta, gemm1_clamp_limit=gemm1_clamp_limit)
    # Blackwell has no int4 tensor-core path and bf16 block scales are not representable as UE8M0 / UE4M3 MMA scale factors:
    # the weights are expanded to bf16 ONCE per weight tensor (cached load-time preparation), the MoE itself is the native pipeline
    w1 = _prepared("mxint4", [gemm1_weights, gemm1_weights_scale], lambda: deq(gemm1_weights, gemm1_weights_scale))
    w2 = _prepared("mxint4", [gemm2_weights, gemm2_weights_scale], lambda: deq(gemm2_weights, gemm2_weights_scale))
    return _into(kw.get("output"), moe_forward(hidden_states, ids, w, w1, w2, local_expert_offset, num_experts))


# ------------------------------------------------------------------ cutlass-style entry point
def cutlass_fused_moe(input: torch.Tensor, token_selected_experts: torch.Tensor, token_final_scales: torch.Tensor,
                      fc1_expert_weights: torch.Tensor, fc2_expert_weights: torch.Tensor, output_dtype: torch.dtype,
                      quant_scales: Optional[List[torch.Tensor]] = None, fc1_expert_biases=None, fc2_expert_biases=None,
                      input_sf=None, swiglu_alpha=None, swiglu_beta=None, swiglu_limit=None, tp_size: int = 1,
                      tp_rank: int = 0, ep_size: int = 1, ep_rank: int = 0, cluster_size: int = 1, cluster_rank: int = 0,
                      output: Optional[torch.Tensor] = None, enable_alltoall: bool = False,
                      use_deepseek_fp8_block_scale: bool = False, use_w4_group_scaling: bool = False,
                      use_mxfp8_act_scaling: bool = False, min_latency_mode: bool = False, use_packed_weights: bool = False,
                      tune_max_num_tokens: int = 8192, enable_pdl=None, activation_type=ActivationType.Swiglu,
                      swizzled_input_sf: bool = True):
    """Pre-routed MoE (reference core.py:775): ``fc1 [E_local, 2I, H]`` (= cat([w3/up, w1/gate])), ``fc2 [E_local, H, I]``.
    ``ep_rank`` selects the local expert range; ``tp_*`` only describe how the caller sharded I."""
    e_local = fc1_expert_weights.shape[0]
    w1, w2 = fc1_expert_weights, fc2_expert_weights
    _reject_unsupported("cutlass_fused_moe", fc1_expert_biases=fc1_expert_biases, fc2_expert_biases=fc2_expert_biases,
                        swiglu_alpha=swiglu_alpha, swiglu_beta=swiglu_beta, swiglu_limit=swiglu_limit,
                        use_w4_group_scaling=use_w4_group_scaling, use_mxfp8_act_scaling=use_mxfp8_act_scaling,
                        min_latency_mode=min_latency_mode, enable_alltoall=enable_alltoall)
    ids, wts = token_selected_experts.int(), token_final_scales.float()
    off, total = ep_rank * e_local, e_local * ep_size
    act = "silu" if int(activation_type) in (int(ActivationType.Swiglu), int(ActivationType.Silu)) else "gelu"
    x = input
    if use_deepseek_fp8_block_scale and quant_scales is not None:
        # [fc1 128x128 block scales, fc2 128x128 block scales]: native fp8 tensor-core pipeline
        xb = x if x.dtype in (torch.float16, torch.bfloat16) else x.to(output_dtype)
        res = moe_forward_fp8_block(xb, None, ids, wts, w1, quant_scales[0], w2, quant_scales[1], off, total, out=output)
    elif w1.dtype == torch.float8_e4m3fn and quant_scales is not None:
        # [fc1_dequant [E], fc2_quant (activation scale), fc2_dequant [E], fc1_input_dequant]: act_q = act * fc2_quant, so the true
        # FC2 de-quantisation scale of a pipeline that quantises activations per group on the fly is fc2_dequant * fc2_quant
        a1 = quant_scales[0].float().reshape(-1)
        a2 = quant_scales[2].float().reshape(-1) * quant_scales[1].float().reshape(-1) if len(quant_scales) > 2 else torch.ones_like(a1)
        res = moe_forward_fp8_per_tensor(x, ids, wts, w1, a1, w2, a2, off, total, out=output)
    elif quant_scales is not None and len(quant_scales) >= 6 and w1.dtype in (torch.uint8, torch.int64, torch.int32):
        # NVFP4: [fc1_act_global, fc1_weight_block, fc1_global, fc2_act_global, fc2_weight_block, fc2_global]; *_global =
        # 1 / (act_global * weight_global): this pipeline quantises activations with a unit global scale, so alpha = global * act_global
        w1q, w2q = w1.view(torch.uint8).reshape(e_local, w1.shape[1], -1), w2.view(torch.uint8).reshape(e_local, w2.shape[1], -1)
        a1 = (quant_scales[2].float() * quant_scales[0].float()).reshape(-1)
        a2 = (quant_scales[5].float() * quant_scales[3].float()).reshape(-1)
        if x.dtype == torch.uint8:                  # NVFP4 activations: block scales 128x4-swizzled by default (reference :916), linear after an
            if input_sf is None:                    # FP4 all-gather / all-to-all (swizzled_input_sf=False)
                raise ValueError("cutlass_fused_moe: NVFP4 input needs input_sf")
            sf = input_sf
            if swizzled_input_sf:
                from ..quantization.fp4 import _unswizzle_index

                kc = x.shape[-1] * 2 // 16
                sf = input_sf.reshape(-1).view(torch.uint8)[_unswizzle_index(x.shape[0], kc).to(input_sf.device)].view(x.shape[0], kc)
            xb = _dequant_nvfp4(x, sf, 1.0 / quant_scales[0].float()).to(output_dtype)
        else:
            xb = x.to(output_dtype)
        res = moe_forward_nvfp4(xb, ids, wts, w1q, quant_scales[1], a1, w2q, quant_scales[4], a2, off, total, out=output)
    else:
        xb = x if x.dtype == output_dtype else x.to(output_dtype)
        res = moe_forward(xb, ids, wts, w1.to(output_dtype), w2.to(output_dtype), local_expert_offset=off, num_experts=total,
                          activation=act, out=output)
    if res.dtype != output_dtype:
        res = res.to(output_dtype)
    return [res]


def cute_dsl_fused_moe_nvfp4(x, x_sf, token_selected_experts, token_final_scales, w1_weight, w1_weight_sf, w1_alpha,
                             fc2_input_scale, w2_weight, w2_weight_sf, w2_alpha, num_experts: int, top_k: int,
                             num_local_experts: Optional[int] = None, local_expert_offset: int = 0, output_dtype=torch.bfloat16,
                             **kw):
    """NVFP4 MoE with pre-computed routing (reference fused_moe/cute_dsl/fused_moe.py) on the native block-scaled grouped tcgen05
    GEMMs (:func:`moe_forward_nvfp4`): gather + activation quantisation, FC1, SwiGLU + re-quantisation, FC2, finalize."""
    xd = _dequant_nvfp4(x, x_sf, 1.0).to(output_dtype) if x.dtype == torch.uint8 else x.to(output_dtype)
    e_local = w1_weight.shape[0]
    ids, wts = token_selected_experts.int(), token_final_scales.float()
    if xd.is_cuda and xd.shape[-1] % 64 == 0 and (w2_weight.shape[2] * 2) % 64 == 0:
        return moe_forward_nvfp4(xd, ids, wts, w1_weight.view(torch.uint8).reshape(e_local, w1_weight.shape[1], -1), w1_weight_sf,
                                 w1_alpha if w1_alpha is not None else 1.0, w2_weight.view(torch.uint8).reshape(e_local, w2_weight.shape[1], -1),
                                 w2_weight_sf, w2_alpha if w2_alpha is not None else 1.0, local_expert_offset, num_experts)
    w1 = _prepared(_scale_tag("nvfp4deq", w1_alpha), [w1_weight, w1_weight_sf, w1_alpha], lambda: _dequant_nvfp4(w1_weight, w1_weight_sf, w1_alpha if w1_alpha is not None else 1.0))
    w2 = _prepared(_scale_tag("nvfp4deq", w2_alpha), [w2_weight, w2_weight_sf, w2_alpha], lambda: _dequant_nvfp4(w2_weight, w2_weight_sf, w2_alpha if w2_alpha is not None else 1.0))
    return moe_forward(xd, ids, wts, w1.to(output_dtype), w2.to(output_dtype), local_expert_offset, num_experts)


class CuteDslMoEWrapper:
    """Stateful wrapper around :func:`cute_dsl_fused_moe_nvfp4` (weights bound once)."""

    def __init__(self, num_experts: int, top_k: int, hidden_size: int, intermediate_size: int,
                 num_local_experts: Optional[int] = None, local_expert_offset: int = 0, **kw) -> None:
        self.num_experts, self.top_k = num_experts, top_k
        self.num_local_experts = num_local_experts or num_experts
        self.local_expert_offset = local_expert_offset

    def run(self, x, x_sf, token_selected_experts, token_final_scales, w1_weight, w1_weight_sf, w1_alpha, fc2_input_scale,
            w2_weight, w2_weight_sf, w2_alpha, **kw):
        return cute_dsl_fused_moe_nvfp4(x, x_sf, token_selected_experts, token_final_scales, w1_weight, w1_weight_sf,
                                        w1_alpha, fc2_input_scale, w2_weight, w2_weight_sf, w2_alpha, self.num_experts,
                                        self.top_k, self.num_local_experts, self.local_expert_offset)


def b12x_fused_moe(*args, **kwargs):
    raise NotImplementedError("b12x_fused_moe targets sm_120/121 GPUs; this library is sm_100a only")


class B12xMoEWrapper:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("B12xMoEWrapper targets sm_120/121 GPUs; this library is sm_100a only")
