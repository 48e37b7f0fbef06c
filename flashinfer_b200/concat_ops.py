"""Parity: reference flashinfer/concat_ops.py:33 (concat_mla_k)."""
from __future__ import annotations

import torch

from . import jit
from .utils import stream_ptr


def concat_mla_k(k: torch.Tensor, k_nope: torch.Tensor, k_rope: torch.Tensor) -> None:
    """``k[..., :nope] = k_nope``; ``k[..., nope:] = k_rope`` broadcast over heads (in place).  Shapes:
    ``k [T, H, nope+rope]``, ``k_nope [T, H, nope]``, ``k_rope [T, 1, rope]``; bf16 / fp16 / fp8."""
    T, H, nope = k_nope.shape
    rope = k_rope.shape[-1]
    if not k.is_cuda:
        k[..., :nope] = k_nope
        k[..., nope:] = k_rope
        return
    if k.stride(-1) != 1 or k_nope.stride(-1) != 1 or k_rope.stride(-1) != 1:
        raise ValueError("concat_mla_k: last dims must be contiguous")
    jit.load("ssm").call("concat_mla_k", k, k_nope, k_rope, T, H, nope, rope, k.stride(0), k.stride(1), k_nope.stride(0),
                         k_nope.stride(1), k_rope.stride(0), k.element_size(), 1, stream_ptr(k))


def get_concat_mla_module(*args, **kwargs):
    """The native module behind this file's ops (reference concat_ops.py get_concat_mla_module: the JIT module accessor)."""
    from . import jit

    return jit.load("ssm")
