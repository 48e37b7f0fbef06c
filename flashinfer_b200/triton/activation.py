"""``flashinfer.triton.activation`` of the reference (a Triton SiLU-and-multiply with optional fp8 scales) on the native activation
kernel (csrc/elementwise/activation.cu).  Without scales and with the input dtype as output dtype it is exactly that kernel; with
scales the de-quantisation / re-quantisation factors are applied around it in fp32."""
from __future__ import annotations

from typing import Optional

import torch

_RANGE = {torch.float8_e4m3fn: 448.0, torch.float8_e5m2: 57344.0, torch.float16: 65504.0, torch.bfloat16: 3.3895313892515355e38}


def scale_and_clamp(x: torch.Tensor, scale, dtype: torch.dtype) -> torch.Tensor:
    """``clamp(x * scale)`` to the finite range of ``dtype``, cast to it (the reference's quantisation epilogue)."""
    hi = _RANGE.get(dtype)
    if hi is None:
        raise TypeError(f"Unsupported dtype: {dtype}")
    s = scale.float() if isinstance(scale, torch.Tensor) else float(scale)
    return (x.float() * s).clamp(-hi, hi).to(dtype)


def silu_and_mul(x: torch.Tensor, x_scale: Optional[torch.Tensor] = None, o_scale: Optional[torch.Tensor] = None,
                 dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """``silu(x[:, :d]) * x[:, d:]`` for ``x [b, 2 d]``.  ``x_scale`` is the scale that was applied to ``x`` (both halves are multiplied
    by it first), ``o_scale`` multiplies the result, which is then clamped to ``dtype``'s range and cast."""
    from ..activation import silu_and_mul as native

    b, n = x.shape
    if n % 2:
        raise ValueError("last dimension must be even")
    o_dtype = dtype or x.dtype
    if x_scale is None and o_scale is None and o_dtype == x.dtype and x.dtype in (torch.float16, torch.bfloat16):
        return native(x)
    xs = x.float() * (x_scale.float() if x_scale is not None else 1.0)
    res = torch.nn.functional.silu(xs[:, : n // 2]) * xs[:, n // 2:]
    return scale_and_clamp(res, o_scale, o_dtype) if o_scale is not None else res.to(o_dtype)
