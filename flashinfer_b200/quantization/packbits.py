"""Bit packing.  Parity: reference flashinfer/quantization/packbits.py:47-139."""
from __future__ import annotations

from typing import Tuple

import torch

from .. import jit
from ..utils import stream_ptr


def packbits(x: torch.Tensor, bitorder: str = "big") -> torch.Tensor:
    """Pack a 1-D bool/uint8 tensor into bits (``bitorder`` 'big' or 'little', numpy.packbits semantics)."""
    if bitorder not in ("big", "little"):
        raise ValueError("bitorder must be 'big' or 'little'")
    n = x.numel()
    x8 = x.reshape(-1).to(torch.uint8)
    out = torch.empty((n + 7) // 8, dtype=torch.uint8, device=x.device)
    if not x.is_cuda:
        pad = (-n) % 8
        b = torch.cat([x8 != 0, torch.zeros(pad, dtype=torch.bool)]).view(-1, 8).to(torch.uint8)
        w = torch.tensor([128, 64, 32, 16, 8, 4, 2, 1] if bitorder == "big" else [1, 2, 4, 8, 16, 32, 64, 128],
                         dtype=torch.uint8)
        out.copy_((b * w).sum(-1).to(torch.uint8))
        return out
    jit.load("quantization").call("packbits", x8.contiguous(), out, None, None, n, 0, 1 if bitorder == "big" else 0,
                                  stream_ptr(x))
    return out


def segment_packbits(x: torch.Tensor, indptr: torch.Tensor, bitorder: str = "big") -> Tuple[torch.Tensor, torch.Tensor]:
    """Pack each segment ``x[indptr[i]:indptr[i+1]]`` separately; returns (packed, new_indptr)."""
    seg_len = (indptr[1:] - indptr[:-1]).long()
    packed_len = (seg_len + 7) // 8
    new_indptr = torch.zeros(indptr.numel(), dtype=indptr.dtype, device=indptr.device)
    new_indptr[1:] = packed_len.cumsum(0)
    total = int(new_indptr[-1])
    out = torch.empty(total, dtype=torch.uint8, device=x.device)
    if not x.is_cuda:
        for i in range(seg_len.numel()):
            out[int(new_indptr[i]) : int(new_indptr[i + 1])] = packbits(x[int(indptr[i]) : int(indptr[i + 1])], bitorder)
        return out, new_indptr
    jit.load("quantization").call("packbits", x.reshape(-1).to(torch.uint8).contiguous(), out, indptr.int().contiguous(),
                                  new_indptr.int().contiguous(), x.numel(), seg_len.numel(),
                                  1 if bitorder == "big" else 0, stream_ptr(x))
    return out, new_indptr
