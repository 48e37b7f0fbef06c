"""FP4 quantisation: NVFP4 (block 16, UE4M3 scale, fp32 global scale) and MXFP4 (block 32, UE8M0 scale).

Parity: reference flashinfer/quantization/fp4_quantization.py:790-1683.  Data are e2m1 pairs packed in uint8
(low nibble = even element); scale factors are uint8 in the 128x4 tile-swizzled layout (the layout
``tcgen05.cp`` moves to TMEM for block-scaled MMA) or linear ``[m, k/vec]``.
"""
from __future__ import annotations

from enum import Enum
from typing import Optional, Tuple

import torch

from .. import jit
from ..utils import dtype_code, round_up, stream_ptr

E2M1_VALUES = [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0]


class SfLayout(Enum):
    layout_128x4 = 0
    layout_8x4 = 1
    layout_linear = 2


def _swizzled_sf_size(m: int, kc: int) -> int:
    return round_up(m, 128) * round_up(kc, 4)


def _unswizzle_index(m: int, kc: int) -> torch.Tensor:
    """flat swizzled offset of every (row, scale-column) pair, row-major ``[m*kc]``."""
    kc_pad = round_up(kc, 4)
    r = torch.arange(m)[:, None]
    c = torch.arange(kc)[None, :]
    tile = (r // 128) * (kc_pad // 4) + c // 4
    return (tile * 512 + (r % 32) * 16 + ((r % 128) // 32) * 4 + c % 4).reshape(-1)


def _index_8x4(m: int, kc: int) -> torch.Tensor:
    """flat offset of every (row, scale-column) pair in the 8x4 tile layout, row-major ``[m*kc]``."""
    kc_pad = round_up(kc, 4)
    r = torch.arange(m)[:, None]
    c = torch.arange(kc)[None, :]
    return (((r // 8) * (kc_pad // 4) + c // 4) * 32 + (r % 8) * 4 + c % 4).reshape(-1)


def _quant_cpu(x: torch.Tensor, gs: float, vec: int, ue8m0: bool):
    m, k = x.shape
    xf = x.float().view(m, k // vec, vec)
    amax = xf.abs().amax(-1)
    if ue8m0:
        from .fp8 import _ue8m0_ceil

        sfb = _ue8m0_ceil(amax / 6.0 * gs)
        sfv = torch.pow(2.0, sfb.float() - 127)
        sfv = torch.where(sfb > 0, sfv, torch.zeros_like(sfv))
    else:
        s8 = (gs * amax / 6.0).to(torch.float8_e4m3fn)
        sfb = s8.view(torch.uint8)
        sfv = s8.float()
    out_scale = torch.where(sfv != 0, gs / sfv, torch.zeros_like(sfv))
    y = xf * out_scale[..., None]
    grid = torch.tensor(E2M1_VALUES)
    mag = y.abs().clamp(max=6.0)
    # round to nearest representable magnitude (ties to even mantissa like cvt.rn)
    idx = torch.bucketize(mag, (grid[1:] + grid[:-1]) / 2)
    mids = (grid[1:] + grid[:-1]) / 2
    tie = (mag[..., None] == mids).any(-1)
    idx = torch.where(tie & (idx % 2 == 1), idx + 1, idx)  # bucketize returns the LOWER code on an exact tie: step up to the even one
    code = idx.to(torch.uint8) | ((y < 0).to(torch.uint8) << 3)
    code = code.view(m, k)
    packed = code[:, 0::2] | (code[:, 1::2] << 4)
    return packed, sfb.reshape(m, k // vec)


def fp4_quantize(input: torch.Tensor, global_scale: Optional[torch.Tensor] = None, sf_vec_size: int = 16,
                 sf_use_ue8m0: bool = False, is_sf_swizzled_layout: bool = True, is_sf_8x4_layout: bool = False,
                 is_global_scale_inversed: bool = False, enable_pdl: Optional[bool] = None,
                 backend: str = "cuda", row_map: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Quantise ``[..., k]`` f16/bf16 to FP4.  Returns (packed uint8 ``[..., k/2]``, uint8 scale factors)."""
    if sf_vec_size not in (16, 32):
        raise ValueError("sf_vec_size must be 16 or 32")
    shape = input.shape
    x = input.reshape(-1, shape[-1])
    m, k = x.shape
    kc = k // sf_vec_size
    gs = global_scale
    if gs is not None and is_global_scale_inversed:
        gs = 1.0 / gs
    sf_size = _swizzled_sf_size(m, kc) if is_sf_swizzled_layout else m * kc
    if is_sf_8x4_layout:  # tiles of 8 rows x 4 scale columns (reference SfLayout.layout_8x4)
        sf_size = round_up(m, 8) * round_up(kc, 4)
    if not x.is_cuda:
        packed, sfb = _quant_cpu(x, float(gs) if gs is not None else 1.0, sf_vec_size, sf_use_ue8m0)
        sf = torch.zeros(sf_size, dtype=torch.uint8)
        if is_sf_8x4_layout:
            sf[_index_8x4(m, kc)] = sfb.reshape(-1)
        elif is_sf_swizzled_layout:
            sf[_unswizzle_index(m, kc)] = sfb.reshape(-1)
        else:
            sf.copy_(sfb.reshape(-1))
    else:
        x = x.contiguous()
        packed = torch.empty(m, k // 2, dtype=torch.uint8, device=x.device)
        sf = torch.zeros(sf_size, dtype=torch.uint8, device=x.device)
        gst = gs.float().reshape(1).contiguous() if gs is not None else None
        jit.load("quantization").call(
            "fp4_quantize", x, packed, sf, gst, 1, m, k, x.stride(0), 0, sf_vec_size, 1 if sf_use_ue8m0 else 0,
            2 if is_sf_8x4_layout else (1 if is_sf_swizzled_layout else 0), 0, row_map, 0, 0, None, 0, 1, dtype_code(x.dtype), 1,
            stream_ptr(x),
        )
    sf = sf.view(-1, round_up(kc, 4)) if (is_sf_swizzled_layout or is_sf_8x4_layout) else sf.view(m, kc)
    return packed.view(*shape[:-1], k // 2), sf


def nvfp4_quantize(a, a_global_sf, sfLayout=SfLayout.layout_128x4, do_shuffle=False, sf_vec_size=16, enable_pdl=None,
                   backend: str = "cuda", *, per_token_activation: bool = False, expanded_idx_to_permuted_idx=None):
    """NVFP4 quantisation with explicit scale-factor layout (reference :1077)."""
    if per_token_activation or expanded_idx_to_permuted_idx is not None:
        raise NotImplementedError("nvfp4_quantize: per-token activation scales / fused MoE row permutation are not implemented "
                                  "(the MoE pipeline quantises inside its gather kernel)")
    q, sf = fp4_quantize(a, a_global_sf, sf_vec_size, False, sfLayout == SfLayout.layout_128x4,
                         sfLayout == SfLayout.layout_8x4)
    if do_shuffle:
        q = shuffle_matrix_a(q, 128)
        sf = shuffle_matrix_sf_a(sf.view(torch.uint8), 128)
    return q, sf


def mxfp4_quantize(a: torch.Tensor, backend: str = "cuda", enable_pdl=None):
    """MXFP4: block 32, UE8M0 scales, swizzled layout (reference :1237)."""
    return fp4_quantize(a, None, 32, True, True)


def nvfp4_batched_quantize(a: torch.Tensor, a_global_sf: torch.Tensor, sf_vec_size: int = 16):
    """``a [b, m, k]`` -> (``[b, m, k/2]`` uint8, ``[b, swizzled]`` scales)."""
    b, m, k = a.shape
    kc = k // sf_vec_size
    per = _swizzled_sf_size(m, kc)
    if not a.is_cuda:
        outs = [fp4_quantize(a[i], a_global_sf, sf_vec_size) for i in range(b)]
        return torch.stack([o[0] for o in outs]), torch.stack([o[1].reshape(-1) for o in outs])
    x = a.contiguous()
    packed = torch.empty(b, m, k // 2, dtype=torch.uint8, device=a.device)
    sf = torch.zeros(b, per, dtype=torch.uint8, device=a.device)
    jit.load("quantization").call(
        "fp4_quantize", x, packed, sf, a_global_sf.float().reshape(1).contiguous(), b, m, k, x.stride(1), x.stride(0),
        sf_vec_size, 0, 1, per, None, 0, 0, None, 0, 1, dtype_code(x.dtype), 1, stream_ptr(x),
    )
    return packed, sf


def scaled_fp4_grouped_quantize(a: torch.Tensor, mask: torch.Tensor, a_global_sf: torch.Tensor):
    """Grouped (per-expert) NVFP4 quantisation of ``a [experts, m, k]`` with one global scale per expert; rows
    beyond ``mask[e]`` are left zero (reference :1512)."""
    e, m, k = a.shape
    outs_q, outs_sf = [], []
    for i in range(e):
        q, sf = fp4_quantize(a[i], a_global_sf[i].reshape(1))
        outs_q.append(q)
        outs_sf.append(sf.reshape(-1))
    return torch.stack(outs_q), torch.stack(outs_sf)


def block_scale_interleave(unswizzled_sf: torch.Tensor) -> torch.Tensor:
    """Linear ``[.., m, kc]`` uint8 scale factors -> 128x4 swizzled layout (flattened per batch)."""
    sf = unswizzled_sf.view(torch.uint8)
    lead = sf.shape[:-2]
    m, kc = sf.shape[-2:]
    b = 1
    for d in lead:
        b *= d
    per = _swizzled_sf_size(m, kc)
    if not sf.is_cuda:
        out = torch.zeros(b, per, dtype=torch.uint8)
        out[:, _unswizzle_index(m, kc)] = sf.reshape(b, -1)
        return out.reshape(-1) if not lead else out.reshape(*lead, per)
    out = torch.zeros(b, per, dtype=torch.uint8, device=sf.device)
    jit.load("quantization").call("sf_interleave", sf.contiguous(), out, b, m, kc, per, 1, stream_ptr(sf))
    return out.reshape(-1) if not lead else out.reshape(*lead, per)


nvfp4_block_scale_interleave = block_scale_interleave


def e2m1_and_ufp8sf_scale_to_float(e2m1_tensor: torch.Tensor, ufp8_scale_tensor: torch.Tensor,
                                   global_scale_tensor: Optional[torch.Tensor] = None, sf_vec_size: int = 16,
                                   ufp8_type: int = 1, is_sf_swizzled_layout: bool = True) -> torch.Tensor:
    """Dequantise packed e2m1 + uint8 scales (ufp8_type 0 = UE8M0, 1 = UE4M3) to fp32 ``[m, k]``."""
    m, k2 = e2m1_tensor.shape
    k = k2 * 2
    kc = k // sf_vec_size
    if e2m1_tensor.is_cuda:
        out = torch.empty(m, k, dtype=torch.float32, device=e2m1_tensor.device)
        gst = global_scale_tensor.float().reshape(1).to(e2m1_tensor.device) if global_scale_tensor is not None else None
        jit.load("quantization").call(
            "fp4_dequantize", e2m1_tensor.contiguous(), ufp8_scale_tensor.view(torch.uint8).contiguous().reshape(-1), gst,
            out, m, k, sf_vec_size, 1 if ufp8_type == 0 else 0, 1 if is_sf_swizzled_layout else 0,
            stream_ptr(e2m1_tensor),
        )
        return out
    sf = ufp8_scale_tensor.view(torch.uint8).reshape(-1)
    if is_sf_swizzled_layout:
        sf = sf[_unswizzle_index(m, kc)]
    sf = sf.view(m, kc)
    scale = torch.pow(2.0, sf.float() - 127) if ufp8_type == 0 else sf.view(torch.float8_e4m3fn).float()
    gs = float(global_scale_tensor) if global_scale_tensor is not None else 1.0
    lut = torch.tensor(E2M1_VALUES + [-v for v in E2M1_VALUES])
    lo, hi = (e2m1_tensor & 0xF).long(), (e2m1_tensor >> 4).long()
    vals = torch.stack([lut[lo], lut[hi]], -1).view(m, k)
    return (vals.view(m, kc, sf_vec_size) * (scale / gs)[..., None]).view(m, k)


def mxfp4_dequantize(a_fp4: torch.Tensor, a_sf: torch.Tensor) -> torch.Tensor:
    return e2m1_and_ufp8sf_scale_to_float(a_fp4, a_sf, None, 32, 0, True)


def mxfp4_dequantize_host(weight: torch.Tensor, scale: torch.Tensor, group_size: int = 32) -> torch.Tensor:
    return e2m1_and_ufp8sf_scale_to_float(weight.cpu(), scale.cpu(), None, group_size, 0, False)


def nvfp4_kv_quantize(input: torch.Tensor, global_scale: torch.Tensor):
    """KV-cache NVFP4 quantisation: ``[..., d]`` -> (packed ``[..., d/2]``, linear UE4M3 scales ``[..., d/16]``)."""
    shape = input.shape
    q, sf = fp4_quantize(input.reshape(-1, shape[-1]), global_scale, 16, False, False)
    return q.view(*shape[:-1], shape[-1] // 2), sf.view(*shape[:-1], shape[-1] // 16)


def nvfp4_kv_dequantize(fp4_data: torch.Tensor, block_scales: torch.Tensor, global_scale: torch.Tensor,
                        output_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    shape = fp4_data.shape
    out = e2m1_and_ufp8sf_scale_to_float(fp4_data.reshape(-1, shape[-1]), block_scales.reshape(-1, shape[-1] // 8),
                                         global_scale, 16, 1, False)
    return out.view(*shape[:-1], shape[-1] * 2).to(output_dtype)


def shuffle_matrix_a(input_tensor: torch.Tensor, epilogue_tile_m: int) -> torch.Tensor:
    """Row permutation used by weight-stationary MMA epilogues: interleave rows so that each group of
    ``epilogue_tile_m`` rows is stored [0, 8, 1, 9, ...]-style (reference :1042)."""
    m = input_tensor.shape[0]
    idx = _shuffle_row_indices(m, epilogue_tile_m, input_tensor.device)
    return input_tensor[idx]


def shuffle_matrix_sf_a(input_tensor: torch.Tensor, epilogue_tile_m: int, num_elts_per_sf: int = 16) -> torch.Tensor:
    """Shuffle the (linear) scale-factor rows the same way as the matrix and re-swizzle (reference :1052)."""
    m = input_tensor.shape[0]
    idx = _shuffle_row_indices(m, epilogue_tile_m, input_tensor.device)
    return block_scale_interleave(input_tensor[idx])


def _shuffle_row_indices(m: int, epilogue_tile_m: int, device) -> torch.Tensor:
    """``idx[new_row] = old_row`` of the trtllm-gen weight shuffle (reference flashinfer/utils.py:801-867): rows are permuted inside
    blocks of 16 rows (32 when ``epilogue_tile_m % 128 == 0``); row ``i`` of a block moves to ``(i % (B/8)) * 8 + i // (B/8)``."""
    b = 32 if epilogue_tile_m % 128 == 0 else 16
    if m % b:
        raise ValueError(f"shuffle_matrix_a: the row count must be a multiple of {b}")
    old = torch.arange(m, device=device)
    i = old % b
    new = (old // b) * b + (i % (b // 8)) * 8 + i // (b // 8)
    idx = torch.empty(m, dtype=torch.long, device=device)
    idx[new] = old
    return idx


def nvfp4_quantize_paged_kv_cache(k_cache: torch.Tensor, v_cache: torch.Tensor, kv_layout: str = "HND",
                                  k_global_sf: Optional[torch.Tensor] = None, v_global_sf: Optional[torch.Tensor] = None):
    """Quantise a bf16 / fp16 paged KV cache to NVFP4 (reference fp4_quantization.py:1365): per-16 UE4M3 block scales in the
    linear cache layout (``head_dim -> head_dim // 16``) plus one fp32 global scale per tensor.
    Returns ``((k_fp4, v_fp4), (k_sf, v_sf), k_global_scale, v_global_scale)``."""
    def one(c: torch.Tensor, gsf):
        if gsf is None:
            gsf = (448.0 * 6.0) / c.float().abs().amax().clamp(min=1e-12)
        gsf = torch.as_tensor(gsf, dtype=torch.float32, device=c.device).reshape(1)
        d = c.shape[-1]
        q, sf = fp4_quantize(c.reshape(-1, d), gsf, 16, False, False)
        return q.view(*c.shape[:-1], d // 2), sf.view(*c.shape[:-1], d // 16).view(torch.float8_e4m3fn), float(1.0 / gsf)

    kq, ksf, kg = one(k_cache, k_global_sf)
    vq, vsf, vg = one(v_cache, v_global_sf)
    return (kq, vq), (ksf, vsf), kg, vg


def nvfp4_dequantize_paged_kv_cache(cache_fp4: torch.Tensor, cache_sf: torch.Tensor, dtype: torch.dtype = torch.bfloat16,
                                    global_scale: Optional[float] = None) -> torch.Tensor:
    """NVFP4 paged KV cache (``[..., head_dim // 2]`` packed e2m1 + ``[..., head_dim // 16]`` UE4M3 block scales, both in the
    cache layout) -> ``dtype`` cache of the same layout.  Used by the attention wrappers when they are handed an NVFP4 cache
    (``kv_cache_sf=``): the block scales are folded in here, the per-tensor global scale rides on ``k_scale`` / ``v_scale``."""
    d2 = cache_fp4.shape[-1]
    flat = cache_fp4.reshape(-1, d2).view(torch.uint8)
    sf = cache_sf.reshape(-1, d2 // 8).view(torch.uint8)
    gs = None if global_scale is None else torch.tensor([1.0 / float(global_scale)], dtype=torch.float32, device=cache_fp4.device)
    out = e2m1_and_ufp8sf_scale_to_float(flat.contiguous(), sf.contiguous(), gs, 16, 1, False)
    return out.to(dtype).view(*cache_fp4.shape[:-1], d2 * 2)


def moe_fp4_quantize(x: torch.Tensor, rows: int, k: int, row_map: torch.Tensor, global_scale: torch.Tensor, gather: bool,
                     gated: bool, sf_vec_size: int = 16, row_list: Optional[torch.Tensor] = None,
                     list_div: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
    """MoE-fused NVFP4 quantiser (one kernel): ``gather`` reads row ``row_map[m]`` of ``x`` (token gather) instead of row
    ``m``; ``gated`` treats a row as ``[linear | gate]`` halves of width ``k`` and quantises ``silu(gate) * linear``
    (SwiGLU fused with the quantisation of the FC2 input).  Rows with ``row_map[m] < 0`` are skipped.  ``row_list``
    (expanded -> permuted row, -1 = not local) makes the kernel visit only the live rows instead of the tile-padded
    matrix (source token of entry ``j`` = ``j // list_div`` when gathering); padding rows / scales stay uninitialised -
    GEMM rows are independent and nothing reads the padding rows of the result.  Returns
    (``[rows, k/2]`` uint8, 128x4-swizzled UE4M3 scales)."""
    kc = k // sf_vec_size
    packed = torch.empty(rows, k // 2, dtype=torch.uint8, device=x.device)
    alloc = torch.empty if row_list is not None else torch.zeros
    sf = alloc(_swizzled_sf_size(rows, kc), dtype=torch.uint8, device=x.device)
    jit.load("quantization").call("fp4_quantize", x, packed, sf, global_scale, 1, rows, k, x.stride(0), 0, sf_vec_size, 0, 1, 0,
                                  row_map, 1 if gather else 0, 1 if gated else 0, row_list,
                                  row_list.numel() if row_list is not None else 0, list_div, dtype_code(x.dtype), 1,
                                  stream_ptr(x))
    return packed, sf.view(-1, round_up(kc, 4))
