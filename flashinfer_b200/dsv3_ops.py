"""DeepSeek-V3 helper ops.  Parity: reference flashinfer/dsv3_ops (mm_M1_16_K7168_N128/N256, fused_topk_deepseek,
concat_mla_k) and flashinfer/gemm/routergemm.py:169-458 (router GEMMs, tinygemm_bf16)."""
from __future__ import annotations

from typing import Optional

import torch

from .concat_ops import concat_mla_k  # noqa: F401
from .fused_moe import fused_topk_deepseek  # noqa: F401
from .gemm.dense import linear


def _router(mat_a: torch.Tensor, mat_b: torch.Tensor, out: Optional[torch.Tensor], k: int, n: int, launch_with_pdl: bool):
    """``mat_a [M<=16, K]`` bf16 x ``mat_b [K, N]`` (column-major, i.e. a ``[N, K]`` weight transposed) -> fp32/bf16 logits.
    Lowered onto the swap-AB + cluster split-K path of the tcgen05 GEMM (the low-latency small-M kernel)."""
    if mat_a.shape[1] != k or mat_b.shape[1] != n:
        raise ValueError(f"router GEMM expects K={k}, N={n}")
    w = mat_b.t()
    if w.stride(1) != 1:
        w = w.contiguous()
    res = linear(mat_a, w, None, enable_pdl=launch_with_pdl)
    if out is not None:
        out.copy_(res)
        return out
    return res


def mm_M1_16_K7168_N256(mat_a, mat_b, out: Optional[torch.Tensor] = None, launch_with_pdl: bool = False):
    return _router(mat_a, mat_b, out, 7168, 256, launch_with_pdl)


def mm_M1_16_K7168_N128(mat_a, mat_b, out: Optional[torch.Tensor] = None, launch_with_pdl: bool = False):
    return _router(mat_a, mat_b, out, 7168, 128, launch_with_pdl)


def mm_M1_16_K6144_N256(mat_a, mat_b, out: Optional[torch.Tensor] = None, launch_with_pdl: bool = False):
    return _router(mat_a, mat_b, out, 6144, 256, launch_with_pdl)


def tinygemm_bf16(input: torch.Tensor, weight: torch.Tensor, out: Optional[torch.Tensor] = None,
                  bias: Optional[torch.Tensor] = None, use_pdl: bool = False) -> torch.Tensor:
    """``out = input @ weight.T + bias`` for tiny M (weight ``[N, K]``)."""
    res = linear(input, weight, bias, enable_pdl=use_pdl)
    if out is not None:
        out.copy_(res)
        return out
    return res
