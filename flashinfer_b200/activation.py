"""Gated activations.  Parity: reference flashinfer/activation.py:77-202."""
from __future__ import annotations

from typing import Optional

import torch

from . import jit, reference
from .utils import device_support_pdl, dtype_code, stream_ptr

_ACT = {"silu": 0, "gelu": 1, "gelu_tanh": 2}


def _act_and_mul(kind: str, input: torch.Tensor, out: Optional[torch.Tensor], enable_pdl, gate_second: bool = False,
                 row_map: Optional[torch.Tensor] = None, row_list: bool = False):
    d = input.shape[-1] // 2
    if input.shape[-1] % 2:
        raise ValueError("last dim must be even")
    if out is None:
        out = torch.empty(*input.shape[:-1], d, dtype=input.dtype, device=input.device)
    if not input.is_cuda:
        if gate_second:
            input = torch.cat([input[..., d:], input[..., :d]], dim=-1)
        if kind == "silu":
            out.copy_(reference.silu_and_mul_ref(input))
        else:
            out.copy_(reference.gelu_and_mul_ref(input, "tanh" if kind == "gelu_tanh" else "none"))
        return out
    x2 = input.reshape(-1, input.shape[-1])
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    o2 = out.view(-1, d)
    pdl = device_support_pdl(input.device) if enable_pdl is None else enable_pdl
    # row_list: row_map is the list of live rows (visit only those) instead of a per-row validity map
    jit.load("activation").call(
        "act_and_mul", x2, o2, row_map.numel() if row_list else x2.shape[0], d, x2.stride(0), o2.stride(0), _ACT[kind],
        1 if gate_second else 0, row_map, 1 if row_list else 0, dtype_code(input.dtype),
        1 if pdl else 0, stream_ptr(input),
    )
    return out


def silu_and_mul(input, out=None, enable_pdl=None):
    """``silu(input[..., :d]) * input[..., d:]``"""
    return _act_and_mul("silu", input, out, enable_pdl)


def gelu_and_mul(input, out=None, enable_pdl=None):
    return _act_and_mul("gelu", input, out, enable_pdl)


def gelu_tanh_and_mul(input, out=None, enable_pdl=None):
    return _act_and_mul("gelu_tanh", input, out, enable_pdl)


def silu_and_mul_scaled_nvfp4_experts_quantize(a: torch.Tensor, mask: torch.Tensor, a_global_sf: torch.Tensor):
    """``a [E, M, 2K]`` -> ``silu(a[..., :K]) * a[..., K:]`` quantised to NVFP4 per expert (rows ``>= mask[e]`` are padding).
    Reference flashinfer/activation.py:204.  Returns ``(fp4 [E, M, K/2] uint8, swizzled scale factors [E, ...])``."""
    from .quantization.fp4 import scaled_fp4_grouped_quantize

    E, M, K2 = a.shape
    act = silu_and_mul(a.reshape(E * M, K2)).view(E, M, K2 // 2)
    gs = a_global_sf.float().reshape(-1)
    if gs.numel() == 1:
        gs = gs.expand(E)
    return scaled_fp4_grouped_quantize(act, mask, gs)


def get_act_and_mul_module(*args, **kwargs):
    """The native module behind this file's ops (reference activation.py get_act_and_mul_module: the JIT module accessor)."""
    from . import jit

    return jit.load("activation")
