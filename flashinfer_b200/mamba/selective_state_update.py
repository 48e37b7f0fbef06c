"""Mamba / Mamba-2 generation-phase selective state update (reference flashinfer/mamba/selective_state_update.py:104).

    dA = exp(A * dt);  state = state * dA + (dt * x) (x) B;  y = state . C + D * x;  y *= z * sigmoid(z)

Single-token ``x [B, (H,) dim]`` and multi-token ``x [B, T, H, dim]`` (speculative verification: the state advances
token by token inside one kernel launch).  Kernel: csrc/elementwise/ssm.cu (one warp per state row, state read and
written exactly once per call, fp32 math).
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import jit
from ..utils import dtype_code, stream_ptr


def selective_state_update_ref(state, x, dt, A, B, C, D=None, z=None, dt_bias=None, dt_softplus=False,
                               state_batch_indices=None, pad_slot_id=-1, update_state=True, dst_state_batch_indices=None):
    """fp32 PyTorch oracle on canonical shapes: state [N,H,dim,ds], x/dt/z [B,T,H,dim], A [H,dim,ds], B/C [B,T,G,ds]."""
    Bsz, T, H, dim = x.shape
    G = B.shape[2]
    out = torch.zeros(Bsz, T, H, dim, dtype=torch.float32, device=x.device)
    for b in range(Bsz):
        slot = int(state_batch_indices[b]) if state_batch_indices is not None else b
        if slot == pad_slot_id:
            continue
        s = state[slot].float().clone()
        for t in range(T):
            dtv = dt[b, t].float() + (dt_bias.float() if dt_bias is not None else 0)
            if dt_softplus:
                dtv = torch.nn.functional.softplus(dtv)
            dA = torch.exp(A.float() * dtv[..., None])
            Bv = B[b, t].float().repeat_interleave(H // G, 0)
            Cv = C[b, t].float().repeat_interleave(H // G, 0)
            s = s * dA + (dtv * x[b, t].float())[..., None] * Bv[:, None, :]
            y = (s * Cv[:, None, :]).sum(-1)
            if D is not None:
                y = y + D.float() * x[b, t].float()
            if z is not None:
                y = y * torch.nn.functional.silu(z[b, t].float())
            out[b, t] = y
        if update_state:
            dslot = int(dst_state_batch_indices[b]) if dst_state_batch_indices is not None else slot
            state[dslot] = s.to(state.dtype)
    return out


def selective_state_update(state: torch.Tensor, x: torch.Tensor, dt: torch.Tensor, A: torch.Tensor, B: torch.Tensor,
                           C: torch.Tensor, D: Optional[torch.Tensor] = None, z: Optional[torch.Tensor] = None,
                           dt_bias: Optional[torch.Tensor] = None, dt_softplus: bool = False,
                           state_batch_indices: Optional[torch.Tensor] = None, pad_slot_id: int = -1,
                           state_scale: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                           disable_state_update: bool = False, intermediate_states_buffer=None, intermediate_state_indices=None,
                           intermediate_state_scales=None, rand_seed=None, philox_rounds: int = 10, cache_steps: int = 0,
                           algorithm: str = "auto", dst_state_batch_indices: Optional[torch.Tensor] = None,
                           cu_seqlens: Optional[torch.Tensor] = None, num_accepted_tokens=None) -> torch.Tensor:
    if (state_scale is not None or intermediate_states_buffer is not None or cu_seqlens is not None or rand_seed is not None
            or intermediate_state_indices is not None or intermediate_state_scales is not None or num_accepted_tokens is not None
            or cache_steps):
        raise NotImplementedError("int16 block-scaled states, intermediate-state caching (buffer / indices / scales / cache_steps / "
                                  "num_accepted_tokens), varlen and stochastic rounding are not implemented")
    # ``algorithm`` selects between the reference's kernel variants (same result) and ``philox_rounds`` only matters with rand_seed
    orig_shape = x.shape
    has_heads = state.dim() == 4
    # canonicalise to state [N,H,dim,ds], x [B,T,H,dim]
    st = state if has_heads else state.unsqueeze(1)
    def canon(t):
        if t is None:
            return None
        if not has_heads:
            t = t.unsqueeze(-2)
        return t if t.dim() == 4 else t.unsqueeze(1)
    xc, dtc, zc = canon(x), canon(dt), canon(z)
    def canon_bc(t):
        if t.dim() == 2:      # [B, ds]
            return t[:, None, None, :]
        if t.dim() == 3:      # [B, G, ds]
            return t[:, None]
        return t
    Bc, Cc = canon_bc(B), canon_bc(C)
    Ac = A if A.dim() == 3 else A.unsqueeze(0)
    Dc = None if D is None else (D if D.dim() == 2 else D.unsqueeze(0))
    dbc = None if dt_bias is None else (dt_bias if dt_bias.dim() == 2 else dt_bias.unsqueeze(0))
    Bsz, T, H, dim = xc.shape
    ds = st.shape[-1]
    G = Bc.shape[2]
    if not x.is_cuda:
        res = selective_state_update_ref(st, xc, dtc.expand(Bsz, T, H, dim), Ac.expand(H, dim, ds), Bc, Cc,
                                         None if Dc is None else Dc.expand(H, dim), zc,
                                         None if dbc is None else dbc.expand(H, dim), dt_softplus, state_batch_indices, pad_slot_id,
                                         not disable_state_update, dst_state_batch_indices).to(x.dtype)
        res = res.reshape(orig_shape)
        if out is not None:
            out.copy_(res)
            return out
        return res
    if xc.stride(-1) != 1:
        xc = xc.contiguous()
    if zc is not None and (zc.stride() != xc.stride()):
        zc = zc.expand_as(xc).contiguous()
        xc = xc.contiguous()
    dtc = dtc.expand(Bsz, T, H, dim).to(xc.dtype) if dtc.dtype != xc.dtype else dtc.expand(Bsz, T, H, dim)
    Af = Ac.float().expand(H, dim, ds)
    if Bc.stride(-1) != 1 or Cc.stride() != Bc.stride():
        Bc, Cc = Bc.contiguous(), Cc.contiguous()
    Bc, Cc = Bc.to(xc.dtype), Cc.to(xc.dtype)
    Df = None if Dc is None else Dc.float().expand(H, dim)
    dbf = None if dbc is None else dbc.float().expand(H, dim)
    if not st.is_contiguous():
        raise ValueError("state must be contiguous")
    res = torch.empty(Bsz, T, H, dim, dtype=xc.dtype, device=x.device)
    strides = torch.tensor([xc.stride(0), xc.stride(1), xc.stride(2), dtc.stride(0), dtc.stride(1), dtc.stride(2), dtc.stride(3),
                            Af.stride(0), Af.stride(1), Af.stride(2), Bc.stride(0), Bc.stride(1), Bc.stride(2),
                            Df.stride(0) if Df is not None else 0, Df.stride(1) if Df is not None else 0,
                            dbf.stride(0) if dbf is not None else 0, dbf.stride(1) if dbf is not None else 0, 0, 0], dtype=torch.int64)
    sidx = state_batch_indices.to(torch.int32).contiguous() if state_batch_indices is not None else None
    didx = dst_state_batch_indices.to(torch.int32).contiguous() if dst_state_batch_indices is not None else None
    jit.load("ssm").call("selective_state_update", st, xc, dtc, Af, Bc, Cc, Df, zc, dbf, res, sidx, didx, strides, Bsz, T, H, dim, ds,
                         G, pad_slot_id, 1 if dt_softplus else 0, 0 if disable_state_update else 1, dtype_code(xc.dtype),
                         dtype_code(st.dtype), 1, stream_ptr(x))
    res = res.reshape(orig_shape)
    if out is not None:
        out.copy_(res)
        return out
    return res
