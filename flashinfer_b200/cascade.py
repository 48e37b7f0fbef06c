"""Attention-state merge operators and cascade wrappers.  Parity: reference flashinfer/cascade.py:42-559."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import jit, reference
from .utils import dtype_code, stream_ptr


def merge_state(v_a, s_a, v_b, s_b) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge two attention states ``(v [n,H,D], s [n,H] base-2 lse)``."""
    if not v_a.is_cuda:
        return reference.merge_state_ref(v_a, s_a, v_b, s_b)
    v_a, v_b = v_a.contiguous(), v_b.contiguous()
    s_a, s_b = s_a.float().contiguous(), s_b.float().contiguous()
    v = torch.empty_like(v_a)
    s = torch.empty_like(s_a)
    n, h, d = v_a.shape
    jit.load("cascade").call("merge_state", v_a, s_a, v_b, s_b, v, s, None, n, h, d, dtype_code(v_a.dtype), 1,
                             stream_ptr(v_a))
    return v, s


def merge_state_in_place(v, s, v_other, s_other, mask: Optional[torch.Tensor] = None) -> None:
    """``(v, s) <- merge((v, s), (v_other, s_other))`` in place; rows with ``mask == False`` are untouched."""
    if not v.is_cuda:
        vo, so = reference.merge_state_ref(v, s, v_other, s_other)
        if mask is not None:
            m = mask.bool()
            v[m] = vo[m]
            s[m] = so[m]
        else:
            v.copy_(vo)
            s.copy_(so)
        return
    if not v.is_contiguous() or not s.is_contiguous() or s.dtype != torch.float32:
        raise ValueError("merge_state_in_place needs contiguous v and fp32 contiguous s")
    n, h, d = v.shape
    m8 = mask.to(torch.uint8).contiguous() if mask is not None else None
    jit.load("cascade").call("merge_state", v, s, v_other.contiguous(), s_other.float().contiguous(), v, s, m8, n, h, d,
                             dtype_code(v.dtype), 1, stream_ptr(v))


def merge_states(v: torch.Tensor, s: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """v [n, K, H, D], s [n, K, H] -> merged ([n, H, D], [n, H])."""
    if not v.is_cuda:
        return reference.merge_states_ref(v, s)
    v = v.contiguous()
    s = s.float().contiguous()
    n, k, h, d = v.shape
    vo = torch.empty(n, h, d, dtype=v.dtype, device=v.device)
    so = torch.empty(n, h, dtype=torch.float32, device=v.device)
    jit.load("cascade").call("merge_states", v, s, vo, so, n, k, h, d, dtype_code(v.dtype), 1, stream_ptr(v))
    return vo, so
