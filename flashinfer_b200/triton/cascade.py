"""``flashinfer.triton.cascade`` of the reference (Triton attention-state merges) on the native merge kernels
(csrc/elementwise/cascade.cu), plus ``variable_length_merge_states`` - a segmented merge with a different number of partial states
per output row (static-shape tensor ops: segment max, exp2 weights, segment sums).  Statistics ``s`` are base-2 log-sum-exps."""
from __future__ import annotations

from typing import Tuple

import torch

from ..cascade import merge_state, merge_state_in_place, merge_states  # noqa: F401  (same semantics as the Triton versions)


def variable_length_merge_states(v: torch.Tensor, s: torch.Tensor, indptr: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """``v [n, H, D]``, ``s [n, H]``: row ``i`` of the result merges the states ``indptr[i] : indptr[i + 1]``
    (``v_out = sum_j 2^(s_j - s_out) v_j``, ``s_out = log2 sum_j 2^s_j``); an empty segment gives zeros and ``-inf``."""
    if v.dim() != 3 or s.dim() != 2 or v.shape[:2] != s.shape:
        raise ValueError("expected v [n, H, D] and s [n, H]")
    n, h, d = v.shape
    rows = indptr.numel() - 1
    ip = indptr.to(torch.int64)
    seg = torch.searchsorted(ip, torch.arange(n, device=v.device), right=True) - 1          # owning output row of every state
    inside = (seg >= 0) & (seg < rows) & (torch.arange(n, device=v.device) < ip[-1])
    seg = seg.clamp(0, max(rows - 1, 0))
    s32 = torch.where(inside[:, None], s.float(), torch.full_like(s, float("-inf"), dtype=torch.float32))
    idx = seg[:, None].expand(n, h)
    top = torch.full((rows, h), float("-inf"), dtype=torch.float32, device=v.device).scatter_reduce(0, idx, s32, "amax", include_self=True)
    w = torch.exp2(s32 - top[seg].nan_to_num(neginf=0.0))                                      # all -inf segments: weights 0
    w = torch.where(torch.isfinite(s32), w, torch.zeros_like(w))
    den = torch.zeros(rows, h, dtype=torch.float32, device=v.device).index_add_(0, seg, w)
    num = torch.zeros(rows, h, d, dtype=torch.float32, device=v.device).index_add_(0, seg, w[..., None] * v.float())
    v_out = (num / den.clamp_min(1e-38)[..., None]).to(v.dtype)
    s_out = torch.where(den > 0, top + torch.log2(den.clamp_min(1e-38)), torch.full_like(den, float("-inf")))
    return v_out, s_out
