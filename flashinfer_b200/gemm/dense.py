"""Dense bf16/fp16 GEMM on the hand-written tcgen05 kernel (csrc/gemm/gemm_bf16_sm100.cu).

``mm_bf16(a, b)`` follows the reference signature (flashinfer/gemm/gemm_base.py:485): ``a`` is
``[m, k]`` row-major and ``b`` is ``[k, n]`` **column-major** (i.e. ``b.T`` is a contiguous
``[n, k]`` weight) — exactly the K-major/K-major "NT" layout the tensor cores want.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from .. import jit
from ..utils import dtype_code, stream_ptr

_workspaces = {}


def _workspace(device) -> torch.Tensor:
    """Per-(device, stream) stream-K workspace: [4 KB self-resetting counters | fp32 partial slots]."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.zeros(32 * 1024 * 1024, dtype=torch.uint8, device=device)  # counters must start at 0
        _workspaces[key] = ws
    return ws


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
           out: Optional[torch.Tensor] = None, enable_pdl: bool = True) -> torch.Tensor:
    """``x[..., K] @ weight[N, K]^T (+ bias[N])`` — the nn.Linear contraction, tcgen05 on CUDA."""
    if weight.dtype != x.dtype or (x.is_cuda and x.dtype not in (torch.float16, torch.bfloat16)):
        raise TypeError("linear: x/weight must both be float16 or bfloat16 (fp32 only on the CPU oracle path)")
    k = x.shape[-1]
    n = weight.shape[0]
    if weight.shape[1] != k:
        raise ValueError(f"linear: K mismatch {weight.shape} vs {x.shape}")
    x2 = x.reshape(-1, k)
    m = x2.shape[0]
    if out is None:
        out = torch.empty(*x.shape[:-1], n, dtype=x.dtype, device=x.device)
    out2 = out.view(-1, n)
    if not x.is_cuda:
        res = x2.float() @ weight.float().t()
        if bias is not None:
            res = res + bias.float()
        out2.copy_(res.to(x.dtype))
        return out
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    if weight.stride(-1) != 1:
        weight = weight.contiguous()
    if _use_cta_pair(m, n, k, bias, x2, weight, out2):
        # large problems: CTA-pair kernel (tcgen05.mma cta_group::2, 256 x 256 tile per SM pair, half the B tile per SM)
        esz = x2.element_size()
        jit.load("gemm_blockscaled_sm100").call(
            "gemm_lowp_nt", x2, weight, out2, None, None, None, None, 1, m, n, k, x2.stride(0) * esz, weight.stride(0) * esz,
            out2.stride(0), 0, 0, 0, 0, 0, 4, 0 if x.dtype == torch.float16 else 1, 0, dtype_code(x.dtype), 0, None, None, None,
            1 if enable_pdl else 0, stream_ptr(x))
        return out
    ws = _workspace(x.device)
    jit.load("gemm_sm100").call(
        "gemm_nt", x2, weight, out2, bias, m, n, k, x2.stride(0), weight.stride(0), out2.stride(0),
        dtype_code(x.dtype), ws, ws.numel(), 1 if enable_pdl else 0, stream_ptr(x),
    )
    return out


_CTA_PAIR = os.environ.get("FIB200_GEMM_2CTA", "1") != "0"


def _use_cta_pair(m, n, k, bias, x2, w, out2) -> bool:
    """Route to the cta_group::2 kernel (csrc/gemm/gemm_blockscaled_sm100.cu, kind 4) when the problem fills the machine with
    256 x 256 pair tiles: per-SM operand traffic per FLOP is half that of the 1-CTA 128 x 256 tile."""
    if not _CTA_PAIR or bias is not None or m < 512 or n < 256 or k < 256:
        return False
    if (x2.stride(0) | w.stride(0) | out2.stride(0)) % 8 or (x2.data_ptr() | w.data_ptr() | out2.data_ptr()) % 16:
        return False
    tiles = ((m + 255) // 256) * ((n + 255) // 256)
    return tiles >= 148  # at least two waves of SM pairs (below that the 1-CTA kernel's finer tiles win)


def interleave_gate_up(w: torch.Tensor) -> torch.Tensor:
    """``[2I, K]`` with the gate rows first and the up rows second -> rows ``(g0, u0, g1, u1, ...)``: the weight layout of
    :func:`linear_gated_silu` (every N tile of the GEMM then holds matching gate / up columns)."""
    half = w.shape[0] // 2
    return torch.stack([w[:half], w[half:]], 1).reshape(w.shape).contiguous()


def linear_gated_silu(x: torch.Tensor, weight_interleaved: torch.Tensor, out: Optional[torch.Tensor] = None,
                      enable_pdl: bool = True) -> torch.Tensor:
    """``silu(x @ Wg.T) * (x @ Wu.T)`` in ONE tcgen05 GEMM: ``weight_interleaved [2I, K]`` from :func:`interleave_gate_up`; the
    SwiGLU runs in the epilogue on the fp32 accumulators and only the ``[M, I]`` result is written (the ``[M, 2I]``
    intermediate and the separate activation kernel disappear).  Equivalent to ``silu_and_mul(linear(x, [Wg; Wu]))``."""
    k = x.shape[-1]
    n2 = weight_interleaved.shape[0]
    x2 = x.reshape(-1, k)
    m = x2.shape[0]
    if out is None:
        out = torch.empty(*x.shape[:-1], n2 // 2, dtype=x.dtype, device=x.device)
    out2 = out.view(-1, n2 // 2)
    if not x.is_cuda:
        h = x2.float() @ weight_interleaved.float().t()
        out2.copy_((torch.nn.functional.silu(h[:, 0::2]) * h[:, 1::2]).to(x.dtype))
        return out
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    w = weight_interleaved if weight_interleaved.stride(-1) == 1 else weight_interleaved.contiguous()
    jit.load("gemm_sm100").call("gemm_nt_gated_silu", x2, w, out2, m, n2, k, x2.stride(0), w.stride(0), out2.stride(0),
                                dtype_code(x.dtype), 1 if enable_pdl else 0, stream_ptr(x))
    return out


def mm_bf16(a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None, pdl: bool = False,
            out: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16, backend: str = "auto"):
    """a [m, k] row-major, b [k, n] column-major -> [m, n]."""
    if b.stride(0) != 1:
        b = b.t().contiguous().t()
    res = linear(a, b.t(), bias, out=out if (out is not None and out.dtype == a.dtype) else None, enable_pdl=pdl)
    if out_dtype != res.dtype:
        res = res.to(out_dtype)
    if out is not None and out.data_ptr() != res.data_ptr():
        out.copy_(res)
        return out
    return res


mm_fp16 = mm_bf16


def bmm_bf16(A: torch.Tensor, B: torch.Tensor, out: Optional[torch.Tensor] = None,
             out_dtype: torch.dtype = torch.bfloat16, backend: str = "auto"):
    """A [B, m, k], B [B, k, n] (column-major per batch) -> [B, m, n]."""
    a, b = A, B
    bsz, m, _ = a.shape
    n = b.shape[-1]
    if out is None:
        out = torch.empty(bsz, m, n, dtype=a.dtype, device=a.device)
    for i in range(bsz):
        mm_bf16(a[i], b[i], out=out[i], out_dtype=a.dtype)
    return out if out_dtype == out.dtype else out.to(out_dtype)


def tgv_gemm_sm100(a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None, pdl: bool = False,
                   out: Optional[torch.Tensor] = None):
    """Low-latency small-M GEMM (reference tgv_gemm_sm100, gemm_base.py:1446): a [m,k], b [k,n] col-major.
    The swap-AB + split-K + PDL path of gemm_nt is the low-latency kernel here."""
    return mm_bf16(a, b, bias=bias, pdl=pdl, out=out, out_dtype=a.dtype)
