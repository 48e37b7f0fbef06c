"""``flashinfer.triton.norm`` of the reference (Triton RMSNorm with optional residual and fp8 scales) on the native norm kernels
(csrc/elementwise/norm.cu): the unscaled forms ARE those kernels; scaled inputs / outputs are handled in fp32 around the same math."""
from __future__ import annotations

from typing import Optional

import torch

from .activation import scale_and_clamp


def _normalise(x32: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    return weight.float() * (x32 * torch.rsqrt((x32 * x32).mean(-1, keepdim=True) + eps))


def rms_norm(x: torch.Tensor, weight: torch.Tensor, out: torch.Tensor, eps: float, in_scale: Optional[torch.Tensor] = None,
             out_scale: Optional[torch.Tensor] = None) -> None:
    """``out[i, j] = x[i, j] * weight[j] / sqrt(eps + mean(x[i]^2))``; ``in_scale`` multiplies ``x`` first, ``out_scale`` multiplies the
    result, which is clamped to ``out``'s dtype range."""
    from ..norm import rmsnorm

    if in_scale is None and out_scale is None and out.dtype == x.dtype and x.dtype in (torch.float16, torch.bfloat16):
        rmsnorm(x, weight, eps, out=out)
        return
    y = _normalise(x.float() * (in_scale.float() if in_scale is not None else 1.0), weight, eps)
    out.copy_(scale_and_clamp(y, out_scale, out.dtype) if out_scale is not None else y.to(out.dtype))


def rms_norm_add_residual(x: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor, eps: float, x_out: Optional[torch.Tensor] = None,
                          x_in_scale: Optional[torch.Tensor] = None, x_out_scale: Optional[torch.Tensor] = None) -> None:
    """``residual += x`` (after ``x_in_scale``), then ``rmsnorm(residual)`` written to ``x_out`` (or back into ``x``)."""
    from ..norm import fused_add_rmsnorm

    if x.shape != residual.shape:
        raise ValueError("x and residual must have the same shape")
    if x_in_scale is None and x_out_scale is None and x_out is None and x.dtype == residual.dtype and x.dtype in (torch.float16, torch.bfloat16):
        fused_add_rmsnorm(x, residual, weight, eps)               # native: in place on both
        return
    r = residual.float() + x.float() * (x_in_scale.float() if x_in_scale is not None else 1.0)
    residual.copy_(r.to(residual.dtype))
    y = _normalise(residual.float(), weight, eps)                  # the stored (rounded) residual is what gets normalised
    dst = x_out if x_out is not None else x
    dst.copy_(scale_and_clamp(y, x_out_scale, dst.dtype) if x_out_scale is not None else y.to(dst.dtype))
