"""SM-constrained GEMM: ``C = alpha * (a @ b) + beta * C`` on a persistent grid of ``num_sms`` CTAs, so that a concurrent kernel
(communication, another stream's work) keeps the remaining SMs.  Parity: reference flashinfer/triton/sm_constraint_gemm.py
(``gemm_persistent`` :14, ``gemm`` :98, ``gemm_descriptor_persistent`` :173) - Triton kernels there, here the persistent tcgen05
GEMM of csrc/gemm/gemm_bf16_sm100.cu with its grid capped through ``gemm_set_sm_budget`` (stream-K remainder and tile order adapt to
the smaller grid)."""
from __future__ import annotations

from typing import Optional

import torch

from .. import jit
from ..utils import device_sm_count, dtype_code, stream_ptr


def _check(a, b, c):
    if a.dim() != 2 or b.dim() != 2:
        raise ValueError("a / b must be 2-D")
    if a.shape[1] != b.shape[0]:
        raise ValueError("Incompatible dimensions between a and b")
    if a.dtype != b.dtype:
        raise ValueError("Incompatible dtypes between a and b")
    if c is not None and (c.dim() != 2 or c.shape != (a.shape[0], b.shape[1])):
        raise ValueError("Incompatible dimensions between a, b and c")


def gemm_persistent(a: torch.Tensor, b: torch.Tensor, c: Optional[torch.Tensor] = None, alpha: float = 1.0, beta: float = 0.0,
                    out_dtype: Optional[torch.dtype] = None, num_sms: Optional[int] = None) -> torch.Tensor:
    """``a [M, K]``, ``b [K, N]`` (any strides; ``w.T`` of a row-major weight is the zero-copy case), ``c [M, N]`` updated in
    place when given.  ``num_sms``: size of the persistent grid (default: all SMs)."""
    _check(a, b, c)
    m, k = a.shape
    n = b.shape[1]
    fp8 = a.dtype in (torch.float8_e4m3fn, torch.float8_e5m2)
    out_dtype = out_dtype or (torch.bfloat16 if fp8 else a.dtype)
    w = b.t()  # [N, K]
    if not a.is_cuda or fp8 or a.dtype not in (torch.float16, torch.bfloat16):
        prod = a.float() @ b.float()
    else:
        if w.stride(-1) != 1:
            w = w.contiguous()
        a2 = a if a.stride(-1) == 1 else a.contiguous()
        prod = torch.empty(m, n, dtype=a.dtype, device=a.device)
        mod = jit.load("gemm_sm100")
        sms = device_sm_count(a.device)
        budget = 0 if num_sms is None else max(1, min(int(num_sms), sms))
        mod.call("gemm_set_sm_budget", budget)
        try:
            mod.call("gemm_nt", a2, w, prod, None, m, n, k, a2.stride(0), w.stride(0), prod.stride(0), dtype_code(a.dtype), None, 0, 0,
                     stream_ptr(a))
        finally:
            mod.call("gemm_set_sm_budget", 0)
    if c is None:
        res = prod if alpha == 1.0 else prod * alpha
        return res.to(out_dtype)
    if beta == 0.0:
        c.copy_((prod if alpha == 1.0 else prod.float() * alpha).to(c.dtype))
    else:
        c.copy_((prod.float() * alpha + c.float() * beta).to(c.dtype))
    return c


def gemm(a, b, c=None, alpha=1.0, beta=0.0, out_dtype=None):
    """Unconstrained flavour (reference :98)."""
    return gemm_persistent(a, b, c, alpha, beta, out_dtype, None)


def gemm_descriptor_persistent(a, b, c=None, alpha=1.0, beta=0.0, out_dtype=None, num_sms=None, EPILOGUE_SUBTILE=False):
    """TMA-descriptor flavour of the reference (:173); the native kernel always loads through TMA descriptors."""
    return gemm_persistent(a, b, c, alpha, beta, out_dtype, num_sms)
