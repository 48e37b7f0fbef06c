"""MXFP8 quantisation (e4m3 data + UE8M0 scale per 32 elements).  Parity: reference
flashinfer/quantization/fp8_quantization.py:163-276."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .. import jit
from ..utils import dtype_code, round_up, stream_ptr
from .fp4 import _swizzled_sf_size, _unswizzle_index


def _ue8m0_ceil(x: torch.Tensor) -> torch.Tensor:
    """smallest power of two >= x, as biased exponent byte."""
    x = x.float().clamp(min=0)
    e = torch.where(x > 0, torch.ceil(torch.log2(x.clamp(min=1e-45))) + 127, torch.zeros_like(x))
    return e.clamp(0, 254).to(torch.uint8)


def mxfp8_quantize(input: torch.Tensor, is_sf_swizzled_layout: bool = True, alignment: int = 32,
                   enable_pdl: Optional[bool] = None, backend: str = "cuda",
                   sf_swizzle_layout=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """``input [m, k]`` (f16/bf16) -> (``[m, k]`` float8_e4m3fn, uint8 UE8M0 scales, 128x4-swizzled or linear)."""
    m, k = input.shape
    if k % 32:
        raise ValueError("mxfp8_quantize: k must be a multiple of 32")
    kc = k // 32
    sf_size = _swizzled_sf_size(m, kc) if is_sf_swizzled_layout else m * kc
    if not input.is_cuda:
        xf = input.float().view(m, kc, 32)
        e = _ue8m0_ceil(xf.abs().amax(-1) / 448.0)
        scale = torch.pow(2.0, e.float() - 127)
        q = (xf / torch.where(scale > 0, scale, torch.ones_like(scale))[..., None]).clamp(-448, 448)
        q = q.view(m, k).to(torch.float8_e4m3fn)
        sf = torch.zeros(sf_size, dtype=torch.uint8)
        if is_sf_swizzled_layout:
            sf[_unswizzle_index(m, kc)] = e.reshape(-1)
        else:
            sf.copy_(e.reshape(-1))
        return q, sf
    x = input.contiguous()
    q = torch.empty(m, k, dtype=torch.float8_e4m3fn, device=x.device)
    sf = torch.zeros(sf_size, dtype=torch.uint8, device=x.device)
    jit.load("quantization").call("mxfp8_quantize", x, q, sf, m, k, x.stride(0), 1 if is_sf_swizzled_layout else 0,
                                  dtype_code(x.dtype), 1, stream_ptr(x))
    return q, sf


def mxfp8_dequantize_host(input: torch.Tensor, scale_tensor: torch.Tensor, is_sf_swizzled_layout: bool = True,
                          sf_swizzle_layout=None) -> torch.Tensor:
    """Reference dequantiser (any device): returns fp32 ``[m, k]``.  ``sf_swizzle_layout`` (``SfLayout.layout_128x4`` or
    ``layout_linear``) overrides ``is_sf_swizzled_layout`` when given."""
    if sf_swizzle_layout is not None:
        from .fp4 import SfLayout

        if sf_swizzle_layout not in (SfLayout.layout_128x4, SfLayout.layout_linear):
            raise ValueError(f"mxfp8_dequantize_host: sf_swizzle_layout must be layout_128x4 or layout_linear, got {sf_swizzle_layout}")
        is_sf_swizzled_layout = sf_swizzle_layout == SfLayout.layout_128x4
    m, k = input.shape
    kc = k // 32
    sf = scale_tensor.reshape(-1)
    if is_sf_swizzled_layout:
        sf = sf[_unswizzle_index(m, kc).to(sf.device)]
    scale = torch.pow(2.0, sf.view(m, kc).float() - 127)
    return (input.float().view(m, kc, 32) * scale[..., None]).view(m, k)
