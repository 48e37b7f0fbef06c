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
    if state_scale is not None or intermediate_state_scales is not None or rand_seed is not None:
        raise NotImplementedError("int16 block-scaled states (state_scale / intermediate_state_scales) and stochastic rounding "
                                  "(rand_seed) are not implemented")
    if intermediate_states_buffer is not None or cu_seqlens is not None or num_accepted_tokens is not None \
            or (state_batch_indices is not None and state_batch_indices.dim() == 2) \
            or (dst_state_batch_indices is not None and dst_state_batch_indices.dim() == 2):
        return _stepwise(state, x, dt, A, B, C, D, z, dt_bias, dt_softplus, state_batch_indices, pad_slot_id, out, disable_state_update,
                         intermediate_states_buffer, intermediate_state_indices, cache_steps, dst_state_batch_indices, cu_seqlens,
                         num_accepted_tokens)
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


def _stepwise(state, x, dt, A, B, C, D, z, dt_bias, dt_softplus, state_batch_indices, pad_slot_id, out, disable_state_update,
              intermediate_states_buffer, intermediate_state_indices, cache_steps, dst_state_batch_indices, cu_seqlens,
              num_accepted_tokens) -> torch.Tensor:
    """Speculative-decoding forms of the update (reference selective_state_update.py :104 - ``intermediate_states_buffer``,
    ``cu_seqlens``, ``num_accepted_tokens``, 2-D state indices): the state has to be observed after every token, so the tokens
    run as rounds of single-token launches on an fp32 working copy of the touched slots - round ``t`` advances every sequence
    that has a ``t``-th token - and each round's state is stored where the caller asked:

    * ``intermediate_states_buffer[intermediate_state_indices[b] (default b), t]`` (multi-token form);
    * with ``num_accepted_tokens``: sequence ``n`` starts from slot ``state_batch_indices[n, max(num_accepted[n] - 1, 0)]`` and the
      state after token ``t`` goes to slot ``dst_state_batch_indices[n, t]`` (``pad_slot_id`` = not stored);
    * otherwise the final state goes to ``dst_state_batch_indices[n]`` (default: the slot it was read from) unless
      ``disable_state_update``.

    The working copy stays fp32 across tokens (stores round to the cache dtype), like the running state of the reference's kernel.
    Sequences whose source slot is ``pad_slot_id`` are skipped: nothing is stored and their output rows are zero."""
    st = state if state.dim() == 4 else state.unsqueeze(1)
    dev = x.device
    varlen = cu_seqlens is not None
    if varlen:
        if x.dim() != 3:
            raise ValueError("cu_seqlens: x must be (total_tokens, nheads, dim)")
        cu = cu_seqlens.tolist()
        flat = lambda t: t                                                            # noqa: E731
    else:
        if x.dim() != 4:
            raise ValueError("intermediate states / 2-D state indices need the multi-token layout x (batch, T, nheads, dim)")
        bsz, T = x.shape[0], x.shape[1]
        cu = [i * T for i in range(bsz + 1)]
        flat = lambda t: None if t is None else t.reshape(bsz * T, *t.shape[2:])     # noqa: E731
    xf, dtf, Bf, Cf, zf = flat(x), flat(dt), flat(B), flat(C), flat(z)
    n = len(cu) - 1
    lens = [cu[i + 1] - cu[i] for i in range(n)]
    max_len = max(lens) if lens else 0
    if cache_steps and max_len > cache_steps:
        raise ValueError(f"a sequence of {max_len} tokens exceeds cache_steps={cache_steps}")
    sbi = state_batch_indices
    if num_accepted_tokens is not None:
        if sbi is None or sbi.dim() != 2:
            raise ValueError("num_accepted_tokens needs 2-D state_batch_indices (N, max_seqlen)")
        init = (num_accepted_tokens.long() - 1).clamp(min=0)
        src = sbi.long()[torch.arange(n, device=sbi.device), init.to(sbi.device)]
    elif sbi is None:
        src = torch.arange(n, device=dev)
    else:
        src = (sbi[:, 0] if sbi.dim() == 2 else sbi).long()
    dst = dst_state_batch_indices if dst_state_batch_indices is not None else sbi
    per_token = num_accepted_tokens is not None
    if per_token:
        dst = dst.long() if dst.dim() == 2 else dst.long().unsqueeze(1)
        if dst.shape[1] < max_len:
            raise ValueError(f"dst_state_batch_indices holds {dst.shape[1]} slots per sequence, the longest sequence has {max_len} tokens")
    else:
        dst = src if dst is None else (dst[:, 0] if dst.dim() == 2 else dst).long()
    src_l = src.tolist()
    live = [i for i in range(n) if src_l[i] != pad_slot_id]
    buf = intermediate_states_buffer
    rows = None
    if buf is not None:
        rows = intermediate_state_indices.long() if intermediate_state_indices is not None else torch.arange(n, device=dev)
        if buf.shape[1] < max_len:
            raise ValueError(f"intermediate_states_buffer caches {buf.shape[1]} steps, the longest sequence has {max_len} tokens")
    res = torch.zeros_like(xf)
    live_t = torch.tensor(live, dtype=torch.long, device=dev)
    work = st[src[live_t].to(st.device)].float().contiguous()                        # [n_live, H, dim, ds]
    pos = {i: j for j, i in enumerate(live)}
    for t in range(max_len):
        act = [i for i in live if lens[i] > t]
        if not act:
            break
        tok = torch.tensor([cu[i] + t for i in act], dtype=torch.long, device=dev)
        sel = torch.tensor([pos[i] for i in act], dtype=torch.long, device=dev)
        whole = len(act) == len(live)
        w = work if whole else work[sel].contiguous()
        res[tok] = selective_state_update(w, xf[tok], dtf[tok], A, Bf[tok], Cf[tok], D, None if zf is None else zf[tok], dt_bias, dt_softplus)
        if not whole:
            work[sel] = w
        act_t = torch.tensor(act, dtype=torch.long, device=dev)
        if buf is not None:
            bw = w if buf.dim() == 5 else w.squeeze(1)
            buf[rows[act_t], t] = bw.to(buf.dtype)
        if per_token:
            slots = dst[act_t.to(dst.device), t]
            keep = slots != pad_slot_id
            st[slots[keep].to(st.device)] = w[keep.to(w.device)].to(st.dtype)
    if not per_token and not disable_state_update and live:
        st[dst[live_t.to(dst.device)].to(st.device)] = work.to(st.dtype)
    res = res.reshape(x.shape)
    if out is not None:
        out.copy_(res)
        return out
    return res
