"""RMSNorm / LayerNorm family.  Parity: reference flashinfer/norm/__init__.py:112-640."""
from __future__ import annotations

from typing import Optional, Union

import torch

from . import jit, reference
from .utils import device_support_pdl, dtype_code, stream_ptr


def _pdl(enable_pdl, t):
    return 1 if (device_support_pdl(t.device) if enable_pdl is None else enable_pdl) else 0


def _strides3(t: torch.Tensor):
    """(rows, heads, s0, s1) for a 2-D [n, d] or 3-D [n, h, d] tensor with contiguous last dim."""
    if t.stride(-1) != 1:
        raise ValueError("last dim must be contiguous")
    if t.ndim == 2:
        return t.shape[0], 1, t.stride(0), 0
    if t.ndim == 3:
        return t.shape[0] * t.shape[1], t.shape[1], t.stride(0), t.stride(1)
    raise ValueError("expected a 2-D or 3-D tensor")


def _rms_launch(x, out, residual, weight, eps, weight_bias, scale, silu, enable_pdl):
    rows, heads, xs0, xs1 = _strides3(x)
    _, _, os0, os1 = _strides3(out)
    rs0 = rs1 = 0
    if residual is not None:
        _, _, rs0, rs1 = _strides3(residual)
    scale_ptr = scale if isinstance(scale, torch.Tensor) else None
    scale_val = 0.0 if isinstance(scale, torch.Tensor) or scale is None else float(scale)
    if scale_ptr is not None and scale_ptr.dtype != torch.float32:
        scale_ptr = scale_ptr.float()
    jit.load("norm").call(
        "rmsnorm_run", x, out, residual, weight, scale_ptr, rows, x.shape[-1], heads, xs0, xs1, os0, os1, rs0, rs1,
        float(eps), float(weight_bias), scale_val, 1 if silu else 0, dtype_code(x.dtype), dtype_code(out.dtype),
        _pdl(enable_pdl, x), stream_ptr(x),
    )


def rmsnorm(input: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6, out: Optional[torch.Tensor] = None,
            enable_pdl: Optional[bool] = None) -> torch.Tensor:
    """``out[i] = input[i] / RMS(input) * weight[i]`` for 2-D (tokens, hidden) or 3-D (tokens, heads, dim) input."""
    if out is None:
        out = torch.empty_like(input)
    if not input.is_cuda:
        out.copy_(reference.rmsnorm_ref(input, weight, eps))
        return out
    _rms_launch(input, out, None, weight, eps, 0.0, None, False, enable_pdl)
    return out


def gemma_rmsnorm(input, weight, eps: float = 1e-6, out=None, enable_pdl=None):
    """Gemma flavour: ``(1 + weight)`` scaling."""
    if out is None:
        out = torch.empty_like(input)
    if not input.is_cuda:
        out.copy_(reference.rmsnorm_ref(input, weight, eps, 1.0))
        return out
    _rms_launch(input, out, None, weight, eps, 1.0, None, False, enable_pdl)
    return out


def fused_add_rmsnorm(input: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6,
                      enable_pdl: Optional[bool] = None) -> None:
    """In place: ``residual += input; input = rmsnorm(residual) * weight``."""
    if not input.is_cuda:
        y, r = reference.fused_add_rmsnorm_ref(input, residual, weight, eps)
        residual.copy_(r)
        input.copy_(y)
        return
    _rms_launch(input, input, residual, weight, eps, 0.0, None, False, enable_pdl)


def gemma_fused_add_rmsnorm(input, residual, weight, eps: float = 1e-6, enable_pdl=None) -> None:
    if not input.is_cuda:
        y, r = reference.fused_add_rmsnorm_ref(input, residual, weight, eps, 1.0)
        residual.copy_(r)
        input.copy_(y)
        return
    _rms_launch(input, input, residual, weight, eps, 1.0, None, False, enable_pdl)


def _quant_ref(y, scale, dtype):
    s = scale.float().item() if isinstance(scale, torch.Tensor) else float(scale)
    lim = 448.0 if dtype == torch.float8_e4m3fn else 57344.0
    return (y.float() / s).clamp(-lim, lim).to(dtype)


def rmsnorm_quant(out: torch.Tensor, input: torch.Tensor, weight: torch.Tensor, scale: Union[float, torch.Tensor],
                  eps: float = 1e-6, enable_pdl: Optional[bool] = None) -> None:
    """``out = (rmsnorm(input) * weight / scale).to(fp8)``."""
    if not input.is_cuda:
        out.copy_(_quant_ref(reference.rmsnorm_ref(input.float(), weight, eps), scale, out.dtype))
        return
    _rms_launch(input, out, None, weight, eps, 0.0, scale, False, enable_pdl)


def fused_add_rmsnorm_quant(out, input, residual, weight, scale, eps: float = 1e-6, enable_pdl=None) -> None:
    """``residual += input; out = (rmsnorm(residual) * weight / scale).to(fp8)``."""
    if not input.is_cuda:
        r = (input.float() + residual.float()).to(input.dtype)
        residual.copy_(r)
        out.copy_(_quant_ref(reference.rmsnorm_ref(r.float(), weight, eps), scale, out.dtype))
        return
    _rms_launch(input, out, residual, weight, eps, 0.0, scale, False, enable_pdl)


def layernorm(input: torch.Tensor, gemma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """LayerNorm with fp32 gamma/beta (argument name ``gemma`` kept for reference parity)."""
    out = torch.empty_like(input)
    if not input.is_cuda:
        out.copy_(reference.layernorm_ref(input, gemma, beta, eps))
        return out
    jit.load("norm").call(
        "layernorm_run", input, out, gemma.float(), beta.float(), input.shape[0], input.shape[-1], input.stride(0),
        out.stride(0), float(eps), dtype_code(input.dtype), 1, stream_ptr(input),
    )
    return out


def fused_rmsnorm_silu(input, weight, eps: float = 1e-6, out=None, block_scale=None):
    """``out = SiLU(RMSNorm(input) * weight)`` (bf16 or fp8 output)."""
    if out is None:
        out = torch.empty_like(input)
    if out.dtype not in (input.dtype, torch.float8_e4m3fn):
        raise NotImplementedError("fused_rmsnorm_silu: nvfp4 output not implemented yet")
    if not input.is_cuda:
        y = reference.rmsnorm_ref(input.float(), weight, eps)
        out.copy_(torch.nn.functional.silu(y).to(out.dtype))
        return out
    _rms_launch(input, out, None, weight, eps, 0.0, None, True, None)
    return out
