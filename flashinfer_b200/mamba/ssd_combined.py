"""Mamba-2 SSD (state-space duality) chunked prefill.  Parity: reference flashinfer/mamba/ssd_combined.py:250 (``SSDCombined``,
a CuTe-DSL tcgen05 kernel there) and its Triton pre-pass ``chunk_cumsum_fwd``.

Algorithm (chunk size Lc, heads H share B / C inside a group):
  dt' = clamp(softplus(dt + bias)),  dA = dt' * A[h],  cs = cumsum(dA) inside each chunk
  intra-chunk :  Y1[l] = sum_{s <= l} (C[l] . B[s]) * exp(cs[l] - cs[s]) * dt'[s] * x[s]          (batched GEMMs + decay mask)
  chunk state :  S_c   = sum_s exp(cs[last] - cs[s]) * dt'[s] * B[s] (x) x[s]                      (batched GEMM)
  recurrence  :  R_c   = exp(cs_{c-1}[last]) * R_{c-1} + S_{c-1},  R_0 = initial state            (sequential over chunks)
  inter-chunk :  Y2[l] = exp(cs[l]) * C[l] . R_c                                                   (batched GEMM)
  y = Y1 + Y2 + D * x;  y *= silu(z)

This implementation composes the four GEMM-shaped stages from batched matmuls (tensor cores through the bf16 / fp32 matmul
path) and keeps the O(#chunks) recurrence on the small state tensors; the single fused tcgen05 kernel of the reference is a
listed gap.  Output layout matches the reference: ``out [B, H, headdim, nchunks, chunk]``,
``final_states [B, H, headdim, dstate]``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def chunk_cumsum_fwd(dt: torch.Tensor, A: torch.Tensor, chunk_size: int, dt_bias: Optional[torch.Tensor] = None,
                     dt_softplus: bool = False, dt_limit: Tuple[float, float] = (0.0, float("inf")),
                     dt_out_dtype: Optional[torch.dtype] = None):
    """``dt [B, L, H]`` -> (``dA_cumsum [B, H, C, Lc]`` fp32, ``dt_processed [B, H, C, Lc]``)."""
    Bsz, L, H = dt.shape
    d = dt.float()
    if dt_bias is not None:
        d = d + dt_bias.float()
    if dt_softplus:
        d = torch.where(d <= 20.0, torch.nn.functional.softplus(d), d)
    d = d.clamp(dt_limit[0], dt_limit[1])
    d = d.view(Bsz, L // chunk_size, chunk_size, H).permute(0, 3, 1, 2)
    dA = d * A.float().view(1, H, 1, 1)
    return dA.cumsum(-1), d.to(dt_out_dtype or torch.float32)


def ssd_combined_fwd(x, dt, A, B, C, chunk_size: int, D=None, z=None, dt_bias=None, dt_softplus: bool = False,
                     dt_limit=(0.0, float("inf")), initial_states=None, seq_idx=None, return_final_states: bool = True,
                     out: Optional[torch.Tensor] = None, compute_dtype: Optional[torch.dtype] = None):
    """``x [B, L, H, P]``, ``dt [B, L, H]``, ``A [H]`` fp32, ``B / C [B, L, G, N]``, ``D [H]`` or ``[H, P]``, ``z`` like ``x``,
    ``initial_states [B (or num_seqs), H, P, N]``; ``seq_idx [B, L]`` marks packed sequences (the state restarts at every
    boundary; ``initial_states`` / final states are then per sequence)."""
    Bsz, L, H, P = x.shape
    G, N = B.shape[2], B.shape[3]
    if L % chunk_size:
        raise ValueError(f"seqlen ({L}) must be divisible by chunk_size ({chunk_size})")
    if seq_idx is not None:
        return _ssd_varlen(x, dt, A, B, C, chunk_size, D, z, dt_bias, dt_softplus, dt_limit, initial_states, seq_idx,
                           return_final_states, out, compute_dtype)
    Cn, Lc = L // chunk_size, chunk_size
    cd = compute_dtype or torch.float32
    cs, dtp = chunk_cumsum_fwd(dt, A, Lc, dt_bias, dt_softplus, dt_limit)              # [B, H, C, Lc]
    rep = H // G
    xc = x.view(Bsz, Cn, Lc, H, P).permute(0, 3, 1, 2, 4)                               # [B, H, C, Lc, P]
    Bc = B.view(Bsz, Cn, Lc, G, N).permute(0, 3, 1, 2, 4)                               # [B, G, C, Lc, N]
    Cc = C.view(Bsz, Cn, Lc, G, N).permute(0, 3, 1, 2, 4)
    # ---- intra-chunk: (C B^T) masked by the decay, applied to dt * x
    CB = torch.matmul(Cc.to(cd), Bc.to(cd).transpose(-1, -2)).float()                   # [B, G, C, Lc, Lc]
    CB = CB.repeat_interleave(rep, 1) if rep > 1 else CB                                # [B, H, C, Lc, Lc]
    seg = cs.unsqueeze(-1) - cs.unsqueeze(-2)                                           # cs[l] - cs[s]
    tri = torch.ones(Lc, Lc, dtype=torch.bool, device=x.device).tril()
    decay = torch.exp(seg.masked_fill(~tri, float("-inf")))
    W = CB * decay * dtp.float().unsqueeze(-2)                                          # weight of source s for target l
    y = torch.matmul(W.to(cd), xc.to(cd)).float()                                       # [B, H, C, Lc, P]
    # ---- chunk states: sum_s exp(cs[last] - cs[s]) * dt[s] * x[s] (x) B[s]
    w_s = torch.exp(cs[..., -1:] - cs) * dtp.float()                                    # [B, H, C, Lc]
    Bh = Bc.repeat_interleave(rep, 1) if rep > 1 else Bc                                # [B, H, C, Lc, N]
    states = torch.matmul((xc.float() * w_s.unsqueeze(-1)).transpose(-1, -2).to(cd), Bh.to(cd)).float()  # [B, H, C, P, N]
    # ---- recurrence over chunks (small tensors): R_c = state entering chunk c
    R = torch.empty(Bsz, H, Cn, P, N, dtype=torch.float32, device=x.device)
    cur = initial_states.float().clone() if initial_states is not None else torch.zeros(Bsz, H, P, N, dtype=torch.float32,
                                                                                          device=x.device)
    chunk_decay = torch.exp(cs[..., -1])                                                # [B, H, C]
    for c in range(Cn):
        R[:, :, c] = cur
        cur = cur * chunk_decay[:, :, c, None, None] + states[:, :, c]
    # ---- inter-chunk: exp(cs[l]) * C[l] . R_c
    Ch = Cc.repeat_interleave(rep, 1) if rep > 1 else Cc                                # [B, H, C, Lc, N]
    y = y + torch.matmul(Ch.to(cd), R.transpose(-1, -2).to(cd)).float() * torch.exp(cs).unsqueeze(-1)
    if D is not None:
        Df = D.float()
        y = y + xc.float() * (Df.view(1, H, 1, 1, -1) if Df.dim() == 2 else Df.view(1, H, 1, 1, 1))
    if z is not None:
        zc = z.view(Bsz, Cn, Lc, H, P).permute(0, 3, 1, 2, 4).float()
        y = y * zc * torch.sigmoid(zc)
    res = y.permute(0, 1, 4, 2, 3)                                                      # [B, H, P, C, Lc]
    if out is not None:
        out.copy_(res)
        res = out
    else:
        res = res.to(x.dtype).contiguous()
    return res, (cur if return_final_states else None)


def _ssd_varlen(x, dt, A, B, C, chunk_size, D, z, dt_bias, dt_softplus, dt_limit, initial_states, seq_idx, return_final_states,
                out, compute_dtype):
    """Packed sequences (batch 1): every sequence is processed on its own, padded to a chunk multiple with dt = -inf-like
    zero steps (dt' = 0 neither decays nor feeds the state)."""
    if x.shape[0] != 1:
        raise ValueError("seq_idx (packed sequences) expects batch == 1")
    L = x.shape[1]
    sid = seq_idx[0].to(torch.int64).cpu()
    bounds = [0] + (torch.nonzero(sid[1:] != sid[:-1]).flatten() + 1).tolist() + [L]
    nseq = len(bounds) - 1
    H, P = x.shape[2], x.shape[3]
    y_full = torch.empty(1, L, H, P, dtype=torch.float32, device=x.device)
    finals = []
    for s in range(nseq):
        lo, hi = bounds[s], bounds[s + 1]
        n = hi - lo
        pad = (-n) % chunk_size

        def cut(t, fill=0.0):
            t = t[:, lo:hi]
            if pad:
                t = torch.cat([t, torch.full((1, pad) + tuple(t.shape[2:]), fill, dtype=t.dtype, device=t.device)], 1)
            return t

        # padded steps: dt' must be exactly 0 -> pre-activate dt here and hand the processed value on
        _, dtp = chunk_cumsum_fwd(dt[:, lo:hi], A, 1, dt_bias, dt_softplus, dt_limit)      # [1, H, n, 1]
        dt_proc = dtp[..., 0].permute(0, 2, 1).contiguous()                                  # [1, n, H], already sliced
        if pad:
            dt_proc = torch.cat([dt_proc, torch.zeros(1, pad, H, dtype=dt_proc.dtype, device=dt_proc.device)], 1)
        init = initial_states[s:s + 1] if initial_states is not None else None
        ys, fin = ssd_combined_fwd(cut(x), dt_proc, A, cut(B), cut(C), chunk_size, D, cut(z) if z is not None else None, None,
                                   False, (0.0, float("inf")), init, None, True, None, compute_dtype)
        y_full[:, lo:hi] = ys.float().permute(0, 3, 4, 1, 2).reshape(1, n + pad, H, P)[:, :n]
        finals.append(fin)
    Cn = L // chunk_size
    res = y_full.view(1, Cn, chunk_size, H, P).permute(0, 3, 4, 1, 2)
    if out is not None:
        out.copy_(res)
        res = out
    else:
        res = res.to(x.dtype).contiguous()
    return res, (torch.cat(finals, 0) if return_final_states else None)


def ssd_reference(x, dt, A, B, C, D=None, z=None, dt_bias=None, dt_softplus=False, initial_states=None):
    """Token-by-token fp32 recurrence (oracle): returns (y [B, L, H, P], final state [B, H, P, N])."""
    Bsz, L, H, P = x.shape
    G, N = B.shape[2], B.shape[3]
    rep = H // G
    s = initial_states.float().clone() if initial_states is not None else torch.zeros(Bsz, H, P, N, device=x.device)
    ys = []
    for t in range(L):
        d = dt[:, t].float() + (dt_bias.float() if dt_bias is not None else 0)
        if dt_softplus:
            d = torch.nn.functional.softplus(d)
        dA = torch.exp(d * A.float())                                                   # [B, H]
        Bt = B[:, t].float().repeat_interleave(rep, 1)
        Ct = C[:, t].float().repeat_interleave(rep, 1)
        s = s * dA[..., None, None] + (d[..., None] * x[:, t].float())[..., None] * Bt[:, :, None, :]
        y = (s * Ct[:, :, None, :]).sum(-1)
        if D is not None:
            y = y + x[:, t].float() * (D.float() if D.dim() == 2 else D.float()[:, None])
        if z is not None:
            zt = z[:, t].float()
            y = y * zt * torch.sigmoid(zt)
        ys.append(y)
    return torch.stack(ys, 1), s


class SSDCombined:
    """Reference-compatible wrapper (flashinfer/mamba/ssd_combined.py:250): construct once per layer geometry, ``run`` per call."""

    def __init__(self, chunk_size: int, nheads: int, headdim: int, dstate: int, ngroups: int,
                 io_dtype: torch.dtype = torch.bfloat16, state_dtype: torch.dtype = torch.bfloat16, has_d: bool = True,
                 d_has_hdim: bool = False, has_initial_states: bool = False, has_varlen: bool = False, has_z: bool = False,
                 seq_idx_dtype=torch.int64) -> None:
        self.chunk_size, self.nheads, self.headdim, self.dstate, self.ngroups = chunk_size, nheads, headdim, dstate, ngroups
        self._state_dtype = state_dtype

    def run(self, x, dt, A, B, C, D=None, z=None, dt_bias=None, dt_softplus: bool = False, dt_limit=(0.0, float("inf")),
            initial_states=None, seq_idx=None, chunk_indices=None, chunk_offsets=None, seq_chunk_cumsum=None,
            update_seq_chunk_cumsum: bool = False, out=None, return_final_states: bool = True):
        if A.dtype != torch.float32:
            raise ValueError(f"A must be float32, got {A.dtype}")
        y, fin = ssd_combined_fwd(x, dt, A, B, C, self.chunk_size, D, z, dt_bias, dt_softplus, dt_limit, initial_states, seq_idx,
                                  return_final_states, out)
        return y, (fin.to(self._state_dtype) if fin is not None else None)
