"""FP8 / MXFP8 / NVFP4 / MXFP4 GEMMs on the block-scaled tcgen05 kernel (csrc/gemm/gemm_blockscaled_sm100.cu).

Parity: reference flashinfer/gemm/gemm_base.py — mm_fp8 (:3792), bmm_fp8 (:6113), mm_mxfp8 (:4621), bmm_mxfp8 (:8311),
mm_fp4 (:5861), gemm_fp8_nt_groupwise (:6280), gemm_fp8_nt_blockscaled (:6632).

Conventions (same as the reference): ``b`` is passed as the column-major ``[k, n]`` view of a ``[n, k]`` weight
(``weight.t()``), block scales are uint8 tensors in the 128x4 swizzled layout produced by ``nvfp4_quantize`` /
``mxfp4_quantize`` / ``mxfp8_quantize`` (2-D linear ``[rows, k/vec]`` scales are swizzled on the fly).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from .. import jit
from ..autotuner import AutoTuner, DynamicTensorSpec, TunableRunner, TuningConfig
from ..quantization.fp4 import _swizzled_sf_size, _unswizzle_index, block_scale_interleave, e2m1_and_ufp8sf_scale_to_float
from ..utils import dtype_code, stream_ptr

_KIND = {"fp8": 0, "mxfp8": 1, "nvfp4": 2, "mxfp4": 3}
_FP8_FMT = {torch.float8_e4m3fn: 0, torch.float8_e5m2: 1}


def _as_nk(b: torch.Tensor) -> torch.Tensor:
    """Accept the reference's column-major ``[.., k, n]`` view and return the K-major ``[.., n, k]`` tensor behind it."""
    bt = b.transpose(-1, -2)
    if bt.stride(-1) != 1:
        bt = bt.contiguous()
    return bt


def _scalar(x, device) -> Optional[torch.Tensor]:
    if x is None:
        return None
    if not isinstance(x, torch.Tensor):
        x = torch.tensor(float(x))
    return x.to(device=device, dtype=torch.float32).reshape(-1)[:1].contiguous()


def _sf_swizzled(sf: torch.Tensor, rows: int, kc: int, batch: int = 1, swizzled: Optional[bool] = None) -> torch.Tensor:
    """Return uint8 ``[batch, round_up(rows,128) * round_up(kc,4)]`` 128x4-swizzled scales.

    ``swizzled=None`` follows the reference's mm_mxfp8 rule: 1-D (or batch x 1-D) tensors are already swizzled,
    ``[.., rows, kc]`` / ``[.., kc, rows]`` tensors are linear and get swizzled here."""
    sf = sf.view(torch.uint8)
    per = _swizzled_sf_size(rows, kc)
    if swizzled is None:
        swizzled = sf.dim() == 1 or (sf.dim() == 2 and sf.shape == (batch, per) and batch > 1)
    if swizzled:
        if sf.numel() != batch * per:
            raise ValueError(f"swizzled scale tensor has {sf.numel()} bytes, expected {batch * per}")
        if not sf.is_contiguous():
            # the reference idiom mm_fp4(a, b.T, a_sf, b_sf.T, ...) hands over a transposed VIEW of the swizzled buffer: the
            # bytes are already in storage order, so undo the view instead of letting reshape() linearise it in logical order
            if sf.dim() >= 2 and sf.transpose(-1, -2).is_contiguous():
                sf = sf.transpose(-1, -2)
            else:
                raise ValueError("swizzled scale tensors must be contiguous (or a plain .T view of a contiguous buffer); "
                                 f"got shape {tuple(sf.shape)} strides {tuple(sf.stride())}")
        return sf.reshape(batch, per)
    if sf.shape[-2:] == (rows, kc):
        return block_scale_interleave(sf.contiguous()).reshape(batch, per)
    if sf.shape[-2:] == (kc, rows):  # transposed linear scales ([k/vec, n])
        return block_scale_interleave(sf.transpose(-1, -2).contiguous()).reshape(batch, per)
    raise ValueError(f"scale tensor of shape {tuple(sf.shape)} is neither swizzled ({per} bytes) nor linear [{rows}, {kc}]")


def _sf_8x4_to_128x4(sf: torch.Tensor, rows: int, kc: int) -> torch.Tensor:
    """Scale bytes stored in 8 x 4 tiles (``SfLayout.layout_8x4``: tile (r / 8, c / 4), 32 bytes, row-major inside) -> linear
    ``[rows, kc]`` uint8 (a gather over rows * kc bytes) -> the 128x4 layout (native ``sf_interleave`` kernel)."""
    from ..quantization.fp4 import _index_8x4

    flat = sf.view(torch.uint8).reshape(-1)
    need = (rows + 7) // 8 * 8 * ((kc + 3) // 4 * 4)
    if flat.numel() < need:
        raise ValueError(f"8x4 scale tensor has {flat.numel()} bytes, expected at least {need} for [{rows}, {kc}] scales")
    return block_scale_interleave(flat[_index_8x4(rows, kc).to(flat.device)].view(rows, kc).contiguous())


def _launch_raw(kind: str, a, b_nk, out, sfa, sfb, alpha_a, alpha_b, K: int, bn: int, tile_expert, meta, row_map):
    B, M, _ = a.shape
    if tile_expert is not None:
        B = b_nk.shape[0]  # grouped: experts
    N = b_nk.shape[1]
    a_fmt = _FP8_FMT.get(a.dtype, 0)
    b_fmt = _FP8_FMT.get(b_nk.dtype, 0)
    jit.load("gemm_blockscaled_sm100").call(
        "gemm_lowp_nt", a, b_nk, out, sfa, sfb, alpha_a, alpha_b, B, M, N, K, a.stride(1), b_nk.stride(1), out.stride(1),
        a.stride(0), b_nk.stride(0), out.stride(0), sfa.stride(0) if sfa is not None else 0,
        sfb.stride(0) if sfb is not None else 0, _KIND[kind], a_fmt, b_fmt, dtype_code(out.dtype), bn, tile_expert, meta,
        row_map, 1, stream_ptr(a))
    return out


class _TileNRunner(TunableRunner):
    """Tactic = the N-tile width of the tcgen05 pipeline (``-1``: the launcher's wave-quantisation heuristic)."""

    def __init__(self, kind, sfa, sfb, alpha_a, alpha_b, K):
        self.args = (kind, sfa, sfb, alpha_a, alpha_b, K)

    def get_valid_tactics(self, inputs, profile):
        n = inputs[1].shape[1]
        return [-1] + [w for w in (64, 128, 192, 256) if w <= max(64, n)]

    def forward(self, inputs, tactic=-1, do_preparation=False, **kwargs):
        kind, sfa, sfb, alpha_a, alpha_b, K = self.args
        a, b_nk, out = inputs
        return _launch_raw(kind, a, b_nk, out, sfa, sfb, alpha_a, alpha_b, K, max(int(tactic), 0), None, None, None)


# the runner carries scale tensors laid out for the live M, so other buckets are tuned when they show up, never synthesised
_TILE_N_CFG = TuningConfig(dynamic_tensor_specs=(DynamicTensorSpec((0, 2), (1, 1)),), use_cold_l2_cache=True, synthesize_buckets=False)


def _launch(kind: str, a: torch.Tensor, b_nk: torch.Tensor, out: torch.Tensor, sfa, sfb, alpha_a, alpha_b, K: int,
            bn: int = 0, tile_expert: Optional[torch.Tensor] = None, meta: Optional[torch.Tensor] = None,
            row_map: Optional[torch.Tensor] = None) -> torch.Tensor:
    """a ``[B, M, Kbytes]``, b_nk ``[B, N, Kbytes]`` (uint8 / fp8 storage, K contiguous), out ``[B, M, N]``.

    ``bn == 0`` lets the launcher pick the N tile; inside ``with autotune():`` (or once tuned configs are loaded) the dense
    path asks the tuner first, keyed by (kind, dtypes, batch, bucketed M, N, K)."""
    bn = bn or int(os.environ.get("FIB200_LOWP_BN", "0"))
    if bn == 0 and tile_expert is None:
        tuner = AutoTuner.get()
        if tuner.is_tuning_mode or tuner.profiling_cache:
            runner = _TileNRunner(kind, sfa, sfb, alpha_a, alpha_b, K)
            _, tactic = tuner.choose_one("gemm_lowp_nt", [runner], _TILE_N_CFG, [a, b_nk, out],
                                         extras=(kind, str(a.dtype), str(out.dtype)))
            bn = max(int(tactic), 0)
    return _launch_raw(kind, a, b_nk, out, sfa, sfb, alpha_a, alpha_b, K, bn, tile_expert, meta, row_map)


def _prep(a: torch.Tensor, b: torch.Tensor):
    a3 = a if a.dim() == 3 else a.unsqueeze(0)
    b3 = b if b.dim() == 3 else b.unsqueeze(0)
    bnk = _as_nk(b3)
    if a3.stride(-1) != 1:
        a3 = a3.contiguous()
    return a3, bnk


def _finish(out3: torch.Tensor, batched: bool, out: Optional[torch.Tensor]):
    res = out3 if batched else out3[0]
    if out is not None and out.data_ptr() != out3.data_ptr():
        out.copy_(res)
        return out
    return res


def _alloc_out(out, shape, dtype, device):
    if out is not None and out.dtype == dtype and out.is_contiguous() and tuple(out.shape) in (shape, shape[1:]):
        return out.reshape(shape)
    return torch.empty(shape, dtype=dtype, device=device)


# ------------------------------------------------------------------ per-tensor fp8
def bmm_fp8(A: torch.Tensor, B: torch.Tensor, A_scale: torch.Tensor, B_scale: torch.Tensor, dtype: torch.dtype,
            out: Optional[torch.Tensor] = None, backend: str = "auto") -> torch.Tensor:
    """``A [b, m, k]`` fp8, ``B [b, k, n]`` fp8 column-major, scalar de-quantisation scales -> ``[b, m, n]``."""
    batched = A.dim() == 3
    a3, bnk = _prep(A, B)
    Bsz, M, K = a3.shape
    N = bnk.shape[1]
    if not A.is_cuda:
        res = (a3.float() @ bnk.float().transpose(-1, -2)) * (A_scale.float() * B_scale.float())
        return _finish(res.to(dtype), batched, out)
    if bnk.shape[0] != Bsz:
        bnk = bnk.expand(Bsz, -1, -1)
    o = _alloc_out(out, (Bsz, M, N), dtype, A.device)
    _launch("fp8", a3, bnk, o, None, None, _scalar(A_scale, A.device), _scalar(B_scale, A.device), K)
    return _finish(o, batched, out)


def mm_fp8(a: torch.Tensor, b: torch.Tensor, alpha: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16,
           out: Optional[torch.Tensor] = None, backend: str = "auto") -> torch.Tensor:
    """``a [m, k]`` fp8 x ``b [k, n]`` fp8 (column-major) * alpha."""
    a3, bnk = _prep(a, b)
    if not a.is_cuda:
        res = a3.float() @ bnk.float().transpose(-1, -2) * (float(alpha) if alpha is not None else 1.0)
        return _finish(res.to(out_dtype), False, out)
    o = _alloc_out(out, (1, a.shape[0], bnk.shape[1]), out_dtype, a.device)
    _launch("fp8", a3, bnk, o, None, None, _scalar(alpha, a.device), None, a.shape[1])
    return _finish(o, False, out)


# ------------------------------------------------------------------ mxfp8
def _mxfp8_ref(a3, bnk, sfa, sfb):
    Bsz, M, K = a3.shape
    N = bnk.shape[1]
    kc = K // 32
    outs = []
    for i in range(Bsz):
        sa = torch.pow(2.0, sfa[i][_unswizzle_index(M, kc).to(sfa.device)].view(M, kc).float() - 127)
        sb = torch.pow(2.0, sfb[i][_unswizzle_index(N, kc).to(sfb.device)].view(N, kc).float() - 127)
        ad = (a3[i].float().view(M, kc, 32) * sa[..., None]).view(M, K)
        bd = (bnk[i].float().reshape(N, kc, 32) * sb[..., None]).view(N, K)
        outs.append(ad @ bd.t())
    return torch.stack(outs)


def bmm_mxfp8(A: torch.Tensor, B: torch.Tensor, A_scale: torch.Tensor, B_scale: torch.Tensor, dtype: torch.dtype,
              out: Optional[torch.Tensor] = None, backend: str = "auto") -> torch.Tensor:
    """MXFP8 (e4m3 data, UE8M0 scale per 32 elements): ``A [b, m, k]``, ``B [b, k, n]`` column-major."""
    batched = A.dim() == 3
    a3, bnk = _prep(A, B)
    Bsz, M, K = a3.shape
    N = bnk.shape[1]
    kc = K // 32
    sfa = _sf_swizzled(A_scale, M, kc, Bsz)
    sfb = _sf_swizzled(B_scale, N, kc, bnk.shape[0])
    if not A.is_cuda:
        return _finish(_mxfp8_ref(a3, bnk, sfa, sfb).to(dtype), batched, out)
    o = _alloc_out(out, (Bsz, M, N), dtype, A.device)
    _launch("mxfp8", a3, bnk, o, sfa, sfb, None, None, K)
    return _finish(o, batched, out)


def mm_mxfp8(a: torch.Tensor, b: torch.Tensor, a_descale: torch.Tensor, b_descale: torch.Tensor,
             out: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16, use_8x4_sf_layout: bool = False,
             backend: str = "auto") -> torch.Tensor:
    if use_8x4_sf_layout:  # activation scales in 8x4 tiles (small-M layout of the reference): re-tiled to the 128x4 layout of tcgen05.cp
        a_descale = _sf_8x4_to_128x4(a_descale, a.shape[-2], a.shape[-1] // 32)
    return bmm_mxfp8(a, b, a_descale, b_descale, out_dtype, out)


# ------------------------------------------------------------------ fp4
def mm_fp4(a: torch.Tensor, b: torch.Tensor, a_descale: torch.Tensor, b_descale: torch.Tensor,
           alpha: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16, out: Optional[torch.Tensor] = None,
           block_size: int = 16, use_8x4_sf_layout: bool = False, backend: str = "auto", use_nvfp4: bool = True,
           enable_pdl: bool = True) -> torch.Tensor:
    """``a [m, k/2]`` packed e2m1, ``b [k/2, n]`` column-major packed e2m1, 128x4-swizzled block scales
    (UE4M3 / 16 for NVFP4, UE8M0 / 32 for MXFP4), ``alpha`` = 1 / (global_sf_a * global_sf_b)."""
    nv = use_nvfp4 and block_size == 16
    vec = 16 if nv else 32
    if use_8x4_sf_layout:  # activation scales in 8x4 tiles (small-M layout of the reference): re-tiled to the 128x4 layout of tcgen05.cp
        a_descale = _sf_8x4_to_128x4(a_descale, a.shape[0], a.shape[1] * 2 // vec)
    a3, bnk = _prep(a.view(torch.uint8), b.view(torch.uint8))
    _, M, K2 = a3.shape
    N = bnk.shape[1]
    K = 2 * K2
    kc = K // vec
    sfa = _sf_swizzled(a_descale, M, kc, 1, True)
    sfb = _sf_swizzled(b_descale, N, kc, 1, True)
    if not a.is_cuda:
        ad = e2m1_and_ufp8sf_scale_to_float(a3[0], sfa[0], None, vec, 1 if nv else 0, True)
        bdq = e2m1_and_ufp8sf_scale_to_float(bnk[0].contiguous(), sfb[0], None, vec, 1 if nv else 0, True)
        res = ad @ bdq.t() * (float(alpha) if alpha is not None else 1.0)
        return _finish(res.to(out_dtype)[None], False, out)
    o = _alloc_out(out, (1, M, N), out_dtype, a.device)
    _launch("nvfp4" if nv else "mxfp4", a3, bnk, o, sfa, sfb, _scalar(alpha, a.device), None, K)
    return _finish(o, False, out)


# ------------------------------------------------------------------ fp8 with fp32 group scales (DeepSeek style)
def _expand_scale(s: torch.Tensor, rows: int, K: int, g_rows: int, g_k: int, major: str) -> torch.Tensor:
    s = s.float()
    if s.shape == (K // g_k, (rows + g_rows - 1) // g_rows) and (major == "MN" or s.shape != ((rows + g_rows - 1) // g_rows, K // g_k)):
        s = s.t()
    return s.repeat_interleave(g_rows, 0)[:rows].repeat_interleave(g_k, 1)[:, :K]


def _scale_strides(s: torch.Tensor, rows_blocks: int, kblocks: int, major: str):
    """(tensor, row stride, k stride) for a scale tensor stored ``[rows, K/128]`` ("K") or ``[K/128, rows]`` ("MN")."""
    s = s.float()
    if s.dim() != 2:
        raise ValueError("scale tensors must be 2-D")
    if s.shape == (rows_blocks, kblocks) and not (major == "MN" and s.shape == (kblocks, rows_blocks)):
        return s, s.stride(0), s.stride(1)
    if s.shape == (kblocks, rows_blocks):
        return s, s.stride(1), s.stride(0)
    raise ValueError(f"scale shape {tuple(s.shape)} does not match ({rows_blocks}, {kblocks})")


def gemm_fp8_nt_groupwise(a: torch.Tensor, b: torch.Tensor, a_scale: torch.Tensor, b_scale: torch.Tensor,
                          scale_major_mode: Optional[str] = None, mma_sm: int = 1,
                          scale_granularity_mnk: Tuple[int, int, int] = (1, 128, 128), out: Optional[torch.Tensor] = None,
                          out_dtype: Optional[torch.dtype] = None, backend: str = "auto") -> torch.Tensor:
    """``a [m, k]`` fp8, ``b [n, k]`` fp8, fp32 scales per (1 x 128) of a and (128 x 128) of b (DeepSeek-V3 recipe).

    Native kernel: every 128-wide K slab goes through the fp8 tensor cores into a fresh TMEM buffer and is promoted into
    fp32 registers with its scale product (csrc/gemm/gemm_blockscaled_sm100.cu ``fp8_groupwise_kernel``).  Other
    granularities fall back to re-scaling into bf16 + the bf16 tcgen05 GEMM."""
    from .dense import mm_bf16

    gm, gn, gk = scale_granularity_mnk
    major = scale_major_mode or "MN"
    M, K = a.shape
    N = b.shape[0]
    out_dtype = out_dtype or (out.dtype if out is not None else torch.bfloat16)
    native = (a.is_cuda and gm == 1 and gn == 128 and gk == 128 and K % 128 == 0 and a.dtype in _FP8_FMT and b.dtype in _FP8_FMT
              and out_dtype in (torch.float16, torch.bfloat16))
    if native:
        kb = K // 128
        sa, sa_row, sa_k = _scale_strides(a_scale, M, kb, major)
        sb, sb_n, sb_k = _scale_strides(b_scale, (N + 127) // 128, kb, major)
        if a.stride(1) != 1:
            a = a.contiguous()
        if b.stride(1) != 1:
            b = b.contiguous()
        res = out if (out is not None and out.dtype == out_dtype and out.stride(-1) == 1) else torch.empty(M, N, dtype=out_dtype, device=a.device)
        jit.load("gemm_blockscaled_sm100").call(
            "gemm_fp8_groupwise_nt", a, b, res, sa, sb, M, N, K, a.stride(0), b.stride(0), res.stride(0), sa_row, sa_k, sb_n, sb_k,
            _FP8_FMT[a.dtype], _FP8_FMT[b.dtype], dtype_code(out_dtype), int(os.environ.get("FIB200_GW_BN", "0")), None, None, 1, 0,
            None, 1, stream_ptr(a))
        if out is not None and out.data_ptr() != res.data_ptr():
            out.copy_(res)
            return out
        return res
    ad = (a.float() * _expand_scale(a_scale, M, K, gm, gk, major)).to(torch.bfloat16)
    bd = (b.float() * _expand_scale(b_scale, N, K, gn, gk, major)).to(torch.bfloat16)
    if not a.is_cuda:
        res = (ad.float() @ bd.float().t()).to(out_dtype)
    else:
        res = mm_bf16(ad, bd.t(), out_dtype=out_dtype)
    if out is not None:
        out.copy_(res)
        return out
    return res


def fp8_group_quantize(x: torch.Tensor, rows: Optional[int] = None, gated: bool = False, row_list: Optional[torch.Tensor] = None,
                       gather: bool = False, list_div: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
    """DeepSeek activation quantisation: ``x [M, K]`` (bf16 / fp16) -> (e4m3 ``[rows, K]``, fp32 scales ``[rows, K / 128]``), one
    scale per 1 x 128 group (``amax / 448``).  MoE modes mirror :func:`flashinfer_b200.quantization.fp4.moe_fp4_quantize`:
    ``row_list`` (expanded -> permuted destination row, only live rows are visited; with ``gather`` the source row of entry
    ``j`` is ``j // list_div``) and ``gated`` (``x`` rows are ``[linear | gate]``, quantise ``silu(gate) * linear``)."""
    K = x.shape[1] // 2 if gated else x.shape[1]
    rows = x.shape[0] if rows is None else rows
    if x.stride(1) != 1:
        x = x.contiguous()
    if not x.is_cuda:
        if row_list is not None:
            raise NotImplementedError("fp8_group_quantize row_list mode is CUDA only")
        v = x.float()
        if gated:
            v = v[:, :K] * torch.nn.functional.silu(v[:, K:])
        g = v.view(rows, K // 128, 128)
        sc = g.abs().amax(-1).clamp_min(1e-10) / 448.0
        return (g / sc[..., None]).view(rows, K).to(torch.float8_e4m3fn), sc
    q = torch.empty(rows, K, dtype=torch.float8_e4m3fn, device=x.device)
    sc = torch.empty(rows, K // 128, dtype=torch.float32, device=x.device)
    jit.load("quantization").call("fp8_group_quantize", x, q, sc, rows, K, x.stride(0), 1 if gated else 0, row_list,
                                  row_list.numel() if row_list is not None else 0, 1 if gather else 0, list_div, dtype_code(x.dtype),
                                  1, stream_ptr(x))
    return q, sc


def grouped_gemm_fp8_groupwise(a: torch.Tensor, a_scale: torch.Tensor, w: torch.Tensor, w_scale: torch.Tensor,
                               tile_expert: torch.Tensor, meta: Optional[torch.Tensor] = None,
                               out_dtype: torch.dtype = torch.bfloat16, out: Optional[torch.Tensor] = None,
                               row_map: Optional[torch.Tensor] = None) -> torch.Tensor:
    """m-grouped contiguous fp8 GEMM with DeepSeek scales, native tcgen05 (``fp8_groupwise_kernel`` grouped mode).

    ``a [M, K]`` e4m3 (rows grouped per expert in 128-row tiles), ``a_scale [M, K/128]`` fp32, ``w [E, N, K]`` e4m3,
    ``w_scale [E, N/128, K/128]`` fp32, ``tile_expert [M/128]`` int32 (-1 = skip the tile), ``meta[0]`` = live row tiles
    (device-side; ``None`` = all).  Reference: DeepGEMM m-grouped contiguous layout (flashinfer/gemm/gemm_base.py
    group_deepgemm_fp8_nt_groupwise) and the fp8 block-scale MoE GEMMs."""
    M, K = a.shape
    E, N, _ = w.shape
    if K % 128 or N % 32 or M % 128:
        raise ValueError("grouped_gemm_fp8_groupwise: K % 128, N % 32 and M % 128 must be 0")
    a = a if a.stride(1) == 1 else a.contiguous()
    w = w.contiguous()
    sa = a_scale.float()
    sb = w_scale if w_scale.dtype == torch.float32 else w_scale.float()  # strides are passed: stride-0 (per-expert scalar) views are fine
    res = out if out is not None else torch.empty(M, N, dtype=out_dtype, device=a.device)
    jit.load("gemm_blockscaled_sm100").call(
        "gemm_fp8_groupwise_nt", a, w, res, sa, sb, M, N, K, a.stride(0), K, res.stride(0), sa.stride(0), sa.stride(1), sb.stride(1),
        sb.stride(2), _FP8_FMT[a.dtype], _FP8_FMT[w.dtype], dtype_code(res.dtype), int(os.environ.get("FIB200_GW_BN", "0")),
        tile_expert, meta, E, sb.stride(0), row_map, 1, stream_ptr(a))
    return res


def gemm_fp8_nt_blockscaled(a, b, a_scale, b_scale, scale_major_mode: Optional[str] = "MN", mma_sm: int = 1,
                            out: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """128 x 128 x 128 block scales on both operands."""
    return gemm_fp8_nt_groupwise(a, b, a_scale, b_scale, scale_major_mode, mma_sm, (128, 128, 128), out, out_dtype)


def fp8_blockscale_gemm_sm90(*args, **kwargs):
    raise NotImplementedError("fp8_blockscale_gemm_sm90 is a Hopper-only entry point; use gemm_fp8_nt_groupwise on B200")


# ------------------------------------------------------------------ trtllm low-latency fp8 GEMM entry points
_LL_CACHE: dict = {}


def prepare_low_latency_gemm_weights(w: torch.Tensor, permutation_indices_cache: Optional[dict] = None) -> torch.Tensor:
    """``w [n, k]`` fp8 -> block layout ``[k // 128, n, 128]`` (reference trtllm_low_latency_gemm.py:199).  No row shuffle is
    applied: the TMA-fed tcgen05 kernel does not need the epilogue-tile permutation of the trtllm-gen cubins."""
    n, k = w.shape
    if k % 128:
        raise ValueError("prepare_low_latency_gemm_weights: k must be a multiple of 128")
    return w.view(n, k // 128, 128).permute(1, 0, 2).contiguous()


def trtllm_low_latency_gemm(A: torch.Tensor, B: torch.Tensor, global_scale: torch.Tensor, out: torch.Tensor) -> None:
    """``out [m, n] = (A [m, k] fp8) x (prepared B [k/128, n, 128] fp8)^T * global_scale`` (small-M fp8 GEMM)."""
    key = (B.data_ptr(), tuple(B.shape))
    w = _LL_CACHE.get(key)
    if w is None:
        kb, n, blk = B.shape
        w = B.permute(1, 0, 2).reshape(n, kb * blk).contiguous()
        if len(_LL_CACHE) > 256:
            _LL_CACHE.clear()
        _LL_CACHE[key] = w
    mm_fp8(A, w.t(), global_scale, out.dtype, out)


def grouped_gemm_nvfp4(a_fp4: torch.Tensor, a_sf: torch.Tensor, w_fp4: torch.Tensor, w_sf: torch.Tensor, alpha: Optional[torch.Tensor],
                       tile_expert: torch.Tensor, meta: Optional[torch.Tensor], out: Optional[torch.Tensor] = None,
                       out_dtype: torch.dtype = torch.bfloat16, vec: int = 16, row_map: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Grouped block-scaled GEMM for MoE (tcgen05 ``kind::mxf4nvf4``): rows of ``a_fp4 [rows, K/2]`` are grouped by expert in
    128-row tiles (``tile_expert[tile]``), ``w_fp4 [E, N, K/2]`` with swizzled scales ``w_sf [E, sf_bytes]`` and one
    ``alpha[e]`` per expert; ``a_sf`` is the swizzled scale tensor of the whole activation matrix.  ``row_map [rows]`` (MoE
    permuted-row -> token, -1 = padding): results of padding rows are not written (they are never read)."""
    rows, K2 = a_fp4.shape
    E, N, _ = w_fp4.shape
    if out is None:
        out = torch.empty(rows, N, dtype=out_dtype, device=a_fp4.device)
    kind = "nvfp4" if vec == 16 else "mxfp4"
    _launch(kind, a_fp4.view(torch.uint8).unsqueeze(0), w_fp4.view(torch.uint8), out.unsqueeze(0), a_sf.view(torch.uint8).reshape(1, -1),
            w_sf.view(torch.uint8).reshape(E, -1), alpha.float().contiguous() if alpha is not None else None, None, 2 * K2,
            tile_expert=tile_expert, meta=meta, row_map=row_map)
    return out
