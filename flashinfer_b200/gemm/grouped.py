"""Grouped / segment GEMMs on the tcgen05 grouped kernel (csrc/gemm/grouped_gemm_sm100.cu).

Parity: reference SegmentGEMMWrapper (flashinfer/gemm/gemm_base.py:1736-1992), grouped_gemm_nt_masked
(gemm/kernels/grouped_gemm_masked_blackwell.py), group_gemm_*_nt_groupwise (:3900-4400), DeepGEMM m-grouped
contiguous / masked layouts (flashinfer/deep_gemm.py:1425-1585), grouped_mm (flashinfer/grouped_mm/core.py).

All variants are lowered onto one layout contract: rows grouped by expert, each group padded to a multiple of
128 rows, ``tile_expert[m_tile]`` naming the expert (-1 = skip) and the live tile count kept on the device.
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import jit
from ..utils import dtype_code, stream_ptr

_TILE = 128


def grouped_gemm_tiles(a: torch.Tensor, w: torch.Tensor, tile_expert: torch.Tensor, meta: Optional[torch.Tensor],
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Low level: ``a [tiles*128, K]`` (row-padded), ``w [E, N, K]`` -> ``out [tiles*128, N]``."""
    rows, K = a.shape
    E, N, _ = w.shape
    if out is None:
        out = torch.empty(rows, N, dtype=a.dtype, device=a.device)
    if not a.is_cuda:
        te = tile_expert.tolist()
        n = int(meta[0]) if meta is not None else len(te)
        for t in range(min(n, len(te))):
            if te[t] >= 0:
                out[t * _TILE:(t + 1) * _TILE] = (a[t * _TILE:(t + 1) * _TILE].float() @ w[te[t]].float().t()).to(a.dtype)
        return out
    jit.load("grouped_gemm_sm100").call("grouped_gemm_nt", a, w.contiguous(), out, tile_expert, meta, rows // _TILE, N, K,
                                        E, a.stride(0), out.stride(0), None, dtype_code(a.dtype), 1, stream_ptr(a))
    return out


def _pad_layout(seg_indptr: torch.Tensor, total: int, weight_indices: Optional[torch.Tensor]):
    """Device-side (sync-free) construction of the 128-padded layout for arbitrary segments."""
    dev = seg_indptr.device
    B = seg_indptr.numel() - 1
    indptr = seg_indptr.to(torch.int64)
    lens = indptr[1:] - indptr[:-1]
    padded = (lens + _TILE - 1) // _TILE * _TILE
    pad_end = torch.cumsum(padded, 0)
    pad_off = pad_end - padded
    rows = torch.arange(total, device=dev, dtype=torch.int64)
    seg = torch.searchsorted(indptr[1:].contiguous(), rows, right=True).clamp_(max=B - 1)
    dest = pad_off[seg] + (rows - indptr[seg])
    max_tiles = (total + B * (_TILE - 1)) // _TILE + 1
    t0 = torch.arange(max_tiles, device=dev, dtype=torch.int64) * _TILE
    tseg = torch.searchsorted(pad_end.contiguous(), t0, right=True)
    valid = tseg < B
    tsegc = tseg.clamp(max=B - 1)
    experts = weight_indices.to(torch.int64)[tsegc] if weight_indices is not None else tsegc
    tile_expert = torch.where(valid, experts, torch.full_like(experts, -1)).to(torch.int32)
    meta = torch.zeros(4, dtype=torch.int32, device=dev)
    meta[0] = (pad_end[-1] // _TILE).to(torch.int32)
    return dest, tile_expert, meta, max_tiles


def segment_gemm(x: torch.Tensor, weights: torch.Tensor, seg_indptr: torch.Tensor,
                 weight_indices: Optional[torch.Tensor] = None, weight_column_major: bool = True,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y[seg i] = x[seg i] @ W[idx(i)]^T  (weights ``[n, N, K]`` when column-major, else ``[n, K, N]``)."""
    if not weight_column_major:
        weights = weights.transpose(1, 2).contiguous()
    total, K = x.shape
    N = weights.shape[1]
    dest, tile_expert, meta, max_tiles = _pad_layout(seg_indptr, total, weight_indices)
    xp = torch.zeros(max_tiles * _TILE, K, dtype=x.dtype, device=x.device)
    xp[dest] = x
    yp = grouped_gemm_tiles(xp, weights, tile_expert, meta)
    y = yp[dest]
    if out is not None:
        out.copy_(y)
        return out
    return y


class SegmentGEMMWrapper:
    """Reference-compatible segment GEMM (LoRA-style batched ``x_i @ W_i``)."""

    def __init__(self, float_workspace_buffer: Optional[torch.Tensor] = None, backend: str = "auto") -> None:
        self._workspace = float_workspace_buffer
        self.backend = "sm100"

    def reset_workspace_buffer(self, float_workspace_buffer: torch.Tensor, int_workspace_buffer=None) -> None:
        self._workspace = float_workspace_buffer

    def run(self, x: torch.Tensor, weights: torch.Tensor, batch_size: int, weight_column_major: bool,
            out: Optional[torch.Tensor] = None, seg_lens: Optional[torch.Tensor] = None,
            seg_indptr: Optional[torch.Tensor] = None, weight_indices: Optional[torch.Tensor] = None) -> torch.Tensor:
        if seg_indptr is None:
            if seg_lens is None:
                raise ValueError("either seg_lens or seg_indptr is required")
            seg_indptr = torch.zeros(batch_size + 1, dtype=torch.int64, device=x.device)
            seg_indptr[1:] = torch.cumsum(seg_lens.to(x.device), 0)
        return segment_gemm(x, weights, seg_indptr.to(x.device), weight_indices, weight_column_major, out)

    forward = run


def grouped_mm_bf16(a: torch.Tensor, b: torch.Tensor, m_indptr: torch.Tensor, out: Optional[torch.Tensor] = None,
                    out_dtype: torch.dtype = torch.bfloat16, *, backend: str = "auto", tactic: int = -1) -> torch.Tensor:
    """``a [cum_m, K]``, ``b [G, N, K]``, ``m_indptr [G+1]`` -> ``[cum_m, N]`` (reference grouped_mm/core.py :82).  ``backend`` /
    ``tactic`` select a cuDNN plan in the reference; there is one native grouped tcgen05 GEMM here."""
    y = segment_gemm(a, b, m_indptr, None, True)
    y = y if y.dtype == out_dtype else y.to(out_dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y


def grouped_mm_fp8(a: torch.Tensor, b: torch.Tensor, m_indptr: torch.Tensor, alpha: Optional[torch.Tensor] = None,
                   out: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16, *, backend: str = "auto",
                   tactic: int = -1) -> torch.Tensor:
    """Per-tensor FP8 grouped GEMM (reference grouped_mm/core.py :204): ``a [cum_m, k]`` / ``b [G, n, k]`` e4m3 or e5m2, ``alpha [1]``
    scales the output.  Composed: fp8 values are exact in bf16, so the operands are widened and the bf16 grouped tcgen05 GEMM
    (fp32 accumulation, bf16 store) runs on the m_indptr segments; ``alpha`` multiplies that result in fp32."""
    if out is not None:
        out_dtype = out.dtype
    y = grouped_mm_bf16(a.to(torch.bfloat16), b.to(torch.bfloat16), m_indptr, None, torch.float32 if alpha is not None else out_dtype)
    if alpha is not None:
        y = (y * alpha.float().reshape(())).to(out_dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y


def grouped_mm_fp4(a: torch.Tensor, b: torch.Tensor, a_descale: torch.Tensor, b_descale: torch.Tensor, m_indptr: torch.Tensor,
                   alpha: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16,
                   block_size: int = 16, *, backend: str = "auto", tactic: int = -1) -> torch.Tensor:
    """Block-scaled FP4 grouped GEMM (reference grouped_mm/core.py :503): ``a [cum_m, k/2]`` / ``b [G, n, k/2]`` packed e2m1 with
    128x4-swizzled block scales ``a_descale [cum_m, k/block_size]`` / ``b_descale [G, n, k/block_size]`` (ue4m3 for
    ``block_size=16`` = NVFP4, ue8m0 for 32 = MXFP4), ``alpha [1]`` scales the output.  Composed: the operands are de-quantised
    (scales un-swizzled) and the bf16 grouped tcgen05 GEMM runs on the m_indptr segments."""
    from ..quantization.fp4 import _swizzled_sf_size, _unswizzle_index

    if block_size not in (16, 32):
        raise ValueError(f"grouped_mm_fp4: block_size must be 16 (NVFP4) or 32 (MXFP4), got {block_size}")
    kind = "ue4m3" if block_size == 16 else "ue8m0"

    def linear_sf(sf, rows, kc):
        flat = sf.reshape(-1).view(torch.uint8)
        if flat.numel() < _swizzled_sf_size(rows, kc):
            raise ValueError(f"block scales hold {flat.numel()} bytes, a 128x4-swizzled [{rows}, {kc}] table needs {_swizzled_sf_size(rows, kc)}")
        return flat[_unswizzle_index(rows, kc).to(flat.device)].view(rows, kc)

    cum_m, k = a.shape[0], a.shape[1] * 2
    G, n = b.shape[0], b.shape[1]
    kc = k // block_size
    ad = _dq_fp4(a, linear_sf(a_descale, cum_m, kc), block_size, kind)
    bd = torch.stack([_dq_fp4(b[g], linear_sf(b_descale[g], n, kc), block_size, kind) for g in range(G)])
    if out is not None:
        out_dtype = out.dtype
    y = grouped_mm_bf16(ad, bd, m_indptr, None, torch.float32 if alpha is not None else out_dtype)
    if alpha is not None:
        y = (y * alpha.float().reshape(())).to(out_dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y


def grouped_mm_mxfp8(a: torch.Tensor, b: torch.Tensor, a_descale: torch.Tensor, b_descale: torch.Tensor, m_indptr: torch.Tensor,
                     out: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16, *, backend: str = "auto",
                     tactic: int = -1) -> torch.Tensor:
    """Grouped MXFP8 GEMM (reference grouped_mm/core.py:346): ``a [cum_m, k]`` e4m3 / e5m2 with 128x4-swizzled UE8M0 scales
    ``[cum_m, k/32]``, ``b [G, n, k]`` with swizzled scales ``[G, n, k/32]``, ``m_indptr [G+1]``.  Composed: the block scales are
    folded in (exact: powers of two) and the bf16 grouped tcgen05 GEMM runs on the m_indptr segments."""
    from ..quantization.fp4 import _swizzled_sf_size, _unswizzle_index

    def dq(x, sf, rows, k):
        kc = k // 32
        sfb = sf.reshape(-1).view(torch.uint8)
        if sfb.numel() >= _swizzled_sf_size(rows, kc) and not (sfb.numel() == rows * kc and rows % 128 == 0 and kc % 4 == 0 and False):
            sfb = sfb[_unswizzle_index(rows, kc).to(sfb.device)]
        scale = torch.exp2(sfb.float().view(rows, kc) - 127.0).repeat_interleave(32, 1)[:, :k]
        return (x.float() * scale).to(torch.bfloat16)

    cum_m, k = a.shape
    G, n, _ = b.shape
    ad = dq(a, a_descale, cum_m, k)
    bd = torch.stack([dq(b[g], b_descale[g], n, k) for g in range(G)])
    return grouped_mm_bf16(ad, bd, m_indptr, out, out_dtype if out is None else out.dtype)


def grouped_gemm_nt_masked(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, masked_m: torch.Tensor) -> torch.Tensor:
    """Masked layout: ``a [E, M_max, K]``, ``b [E, N, K]``, only the first ``masked_m[e]`` rows per expert are
    computed (tiles past the mask are skipped, rows inside a partially-masked tile are still written)."""
    E, M, K = a.shape
    if M % _TILE:
        raise ValueError("grouped_gemm_nt_masked: M_max must be a multiple of 128")
    tpe = M // _TILE
    t = torch.arange(E * tpe, device=a.device)
    e = t // tpe
    valid = (t % tpe) * _TILE < masked_m.to(a.device)[e]
    tile_expert = torch.where(valid, e, torch.full_like(e, -1)).to(torch.int32)
    grouped_gemm_tiles(a.reshape(E * M, K), b, tile_expert, None, out.reshape(E * M, -1))
    return out


# ---- quantised grouped GEMMs: de-quantise to bf16, then the tcgen05 grouped kernel ---------------------------------
def _dq_groupwise(x: torch.Tensor, scale: torch.Tensor, gran_rows: int, gran_k: int, scale_major_k: bool) -> torch.Tensor:
    s = scale.float()
    if not scale_major_k:
        s = s.transpose(-1, -2)
    s = s.repeat_interleave(gran_rows, -2).repeat_interleave(gran_k, -1)[..., : x.shape[-2], : x.shape[-1]]
    return (x.float() * s).to(torch.bfloat16)


def group_gemm_fp8_nt_groupwise(a: torch.Tensor, b: torch.Tensor, a_scale: torch.Tensor, b_scale: torch.Tensor,
                                m_indptr: torch.Tensor, scale_granularity_mnk=(1, 128, 128), scale_major_mode: str = "MN",
                                mma_sm: int = 1, out: Optional[torch.Tensor] = None,
                                out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    gm, gn, gk = scale_granularity_mnk
    major_k = scale_major_mode == "K"
    ad = _dq_groupwise(a, a_scale, gm, gk, major_k)
    bd = _dq_groupwise(b, b_scale, gn, gk, major_k)
    return grouped_mm_bf16(ad, bd, m_indptr, out, out_dtype)


def _native_fp8_groupwise(a, b, a_scale, b_scale, gran) -> bool:
    return (a.is_cuda and tuple(gran) == (1, 128, 128) and a.dtype == torch.float8_e4m3fn and b.dtype == torch.float8_e4m3fn
            and a.shape[-1] % 128 == 0 and b.shape[-2] % 32 == 0 and a_scale.shape[-1] == a.shape[-1] // 128
            and b_scale.shape[-1] == a.shape[-1] // 128 and b_scale.dim() == 3)


def group_deepgemm_fp8_nt_groupwise(a, b, a_scale, b_scale, m_indices: torch.Tensor, scale_granularity_mnk=(1, 128, 128),
                                    out: Optional[torch.Tensor] = None, out_dtype=torch.bfloat16) -> torch.Tensor:
    """DeepGEMM m-grouped contiguous: ``m_indices [M]`` names the group of every row (groups 128-aligned).  With the
    DeepSeek (1, 128, 128) granularity this is ONE native fp8 tcgen05 kernel (per-slab TMEM promotion with the scale
    product, ``gemm_fp8_groupwise_nt`` grouped mode); other granularities de-quantise into the bf16 grouped GEMM."""
    gm, gn, gk = scale_granularity_mnk
    tile_expert = m_indices[::_TILE].to(torch.int32).contiguous()
    if _native_fp8_groupwise(a, b, a_scale, b_scale, scale_granularity_mnk) and a.shape[0] % _TILE == 0 and a_scale.dim() == 2:
        from .lowp import grouped_gemm_fp8_groupwise

        odt = out_dtype if out_dtype in (torch.float16, torch.bfloat16) else torch.bfloat16
        y = grouped_gemm_fp8_groupwise(a, a_scale, b, b_scale, tile_expert, None, odt,
                                       out if (out is not None and out.dtype == odt and out.is_contiguous()) else None)
    else:
        ad = _dq_groupwise(a, a_scale, gm, gk, True)
        bd = _dq_groupwise(b, b_scale, gn, gk, True)
        y = grouped_gemm_tiles(ad, bd, tile_expert, None)
    if out is not None and y.data_ptr() != out.data_ptr():
        out.copy_(y)
        return out
    return y if y.dtype == out_dtype else y.to(out_dtype)


def batch_deepgemm_fp8_nt_groupwise(a, b, a_scale, b_scale, masked_m: torch.Tensor, expected_m: int = 0,
                                    scale_granularity_mnk=(1, 128, 128), out: Optional[torch.Tensor] = None,
                                    out_dtype=torch.bfloat16) -> torch.Tensor:
    """DeepGEMM masked layout: ``a [G, M_max, K]`` fp8 + ``a_scale [G, M_max, K/128]``; rows ``>= masked_m[g]`` are padding.
    Native for (1, 128, 128) scales and ``M_max % 128 == 0``: the batch is one m-grouped problem whose tile -> expert map
    marks the tiles past ``masked_m`` as skipped (built on the device, no host sync)."""
    gm, gn, gk = scale_granularity_mnk
    G, Mx, K = a.shape
    if _native_fp8_groupwise(a, b, a_scale, b_scale, scale_granularity_mnk) and Mx % _TILE == 0 and a.is_contiguous():
        from .lowp import grouped_gemm_fp8_groupwise

        mt = Mx // _TILE
        t = torch.arange(G * mt, device=a.device, dtype=torch.int32)
        g = t // mt
        te = torch.where((t % mt) * _TILE < masked_m.to(torch.int32)[g.long()], g, torch.full_like(g, -1)).contiguous()
        odt = out_dtype if out_dtype in (torch.float16, torch.bfloat16) else torch.bfloat16
        o = grouped_gemm_fp8_groupwise(a.view(G * Mx, K), a_scale.reshape(G * Mx, K // 128), b, b_scale, te, None, odt)
        o = o.view(G, Mx, b.shape[1])
    else:
        ad = _dq_groupwise(a, a_scale, gm, gk, True)
        bd = _dq_groupwise(b, b_scale, gn, gk, True)
        o = torch.empty(a.shape[0], a.shape[1], b.shape[1], dtype=torch.bfloat16, device=a.device)
        grouped_gemm_nt_masked(ad, bd, o, masked_m)
    o = o if o.dtype == out_dtype else o.to(out_dtype)
    if out is not None:
        out.copy_(o)
        return out
    return o


def _dq_fp4(x: torch.Tensor, sf: torch.Tensor, vec: int, sf_dtype) -> torch.Tensor:
    from ..quantization.fp4 import E2M1_VALUES

    lut = torch.tensor(E2M1_VALUES + [-v for v in E2M1_VALUES], device=x.device)
    xb = x.view(torch.uint8)
    vals = torch.stack([lut[(xb & 0xF).long()], lut[(xb >> 4).long()]], -1).flatten(-2)
    sb = sf.view(torch.uint8)
    s = sb.view(torch.float8_e4m3fn).float() if sf_dtype == "ue4m3" else torch.exp2(sb.float() - 127.0)
    s = s.reshape(*vals.shape[:-1], -1).repeat_interleave(vec, -1)[..., : vals.shape[-1]]
    return (vals * s).to(torch.bfloat16)


def group_gemm_mxfp4_nt_groupwise(a, b, a_scale, b_scale, m_indptr, mma_sm: int = 1, tile_m: int = 128, tile_n: int = 128,
                                  tile_k: int = 128, swap_ab: bool = True, out=None, out_dtype=torch.bfloat16):
    """a: mxfp8 e4m3 ``[cum_m, K]`` + ue8m0 ``[cum_m, K/32]``; b: mxfp4 ``[G, N, K/2]`` + ue8m0 ``[G, N, K/32]`` (linear SF)."""
    s = torch.exp2(a_scale.view(torch.uint8).float() - 127.0).repeat_interleave(32, -1)[:, : a.shape[1]]
    ad = (a.float() * s).to(torch.bfloat16)
    bd = _dq_fp4(b, b_scale, 32, "ue8m0")
    return grouped_mm_bf16(ad, bd, m_indptr, out, out_dtype)


def group_gemm_nvfp4_nt_groupwise(a, b, a_scale, b_scale, m_indptr, alpha: Optional[torch.Tensor] = None, out=None,
                                  out_dtype=torch.bfloat16, **kw):
    """a: nvfp4 ``[cum_m, K/2]`` + ue4m3 ``[cum_m, K/16]``; b: nvfp4 ``[G, N, K/2]`` + ue4m3 ``[G, N, K/16]`` (linear SF);
    ``alpha [G]`` = per-group global scale product."""
    ad = _dq_fp4(a, a_scale, 16, "ue4m3")
    bd = _dq_fp4(b, b_scale, 16, "ue4m3")
    if alpha is not None:
        bd = (bd.float() * alpha.float().reshape(-1, 1, 1)).to(torch.bfloat16)
    return grouped_mm_bf16(ad, bd, m_indptr, out, out_dtype)
