"""Gated DeltaNet (gated delta rule linear attention): decode step, multi-token verification and prefill.

Parity: reference flashinfer/gdn_decode.py (gated_delta_rule_decode :416, _pretranspose :118, gated_delta_rule_mtp :557)
and flashinfer/gdn_prefill.py (chunk_gated_delta_rule :100).  Recurrence per (sequence, value head), state S [K, V]:

    S <- exp(g_t) * S;   v' = beta_t * (v_t - k_t^T S);   S <- S + k_t (x) v';   o_t = scale * q_t^T S

Kernel: csrc/elementwise/ssm.cu ``gated_delta_rule`` — one CTA per (sequence, v-head) keeps its state column-block in
registers for the whole token loop, so the state crosses HBM exactly twice per call regardless of T.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple, Union

import torch

from . import jit
from .utils import dtype_code, stream_ptr


def gated_delta_rule_ref(q, k, v, state, g_log, beta, scale, l2norm):
    """fp32 oracle: q/k [B,T,H,K], v [B,T,HV,V], state [B,HV,K,V] (updated in place), g_log/beta [B,T,HV]."""
    B, T, H, K = q.shape
    HV, V = v.shape[2], v.shape[3]
    out = torch.zeros(B, T, HV, V, dtype=torch.float32, device=q.device)
    rep = HV // H
    for b in range(B):
        S = state[b].float().clone()
        for t in range(T):
            qt = q[b, t].float().repeat_interleave(rep, 0)
            kt = k[b, t].float().repeat_interleave(rep, 0)
            if l2norm:
                qt = qt * torch.rsqrt((qt * qt).sum(-1, keepdim=True) + 1e-6)
                kt = kt * torch.rsqrt((kt * kt).sum(-1, keepdim=True) + 1e-6)
            S = S * torch.exp(g_log[b, t].float())[:, None, None]
            ks = torch.einsum("hk,hkv->hv", kt, S)
            vn = (v[b, t].float() - ks) * beta[b, t].float()[:, None]
            S = S + kt[:, :, None] * vn[:, None, :]
            out[b, t] = torch.einsum("hk,hkv->hv", qt * scale, S)
        state[b] = S.to(state.dtype)
    return out


def gated_delta_rule_chunked(q, k, v, state, g_log, beta, scale, l2norm, chunk: int = 64, compute_dtype: torch.dtype = torch.float32):
    """Chunk-parallel gated delta rule (WY form), composed of batched GEMMs + one triangular solve per chunk - the algorithm of
    the reference's tcgen05 chunked-prefill kernel (gdn_kernels/blackwell/gated_delta_net_chunked.py), here as tensor ops.

    Same contract as :func:`gated_delta_rule_ref` (``state`` is updated in place).  Inside a chunk with entering state ``S0``,
    cumulative log-gates ``G_i`` and ``Gamma_i = exp(G_i)``::

        (I + A) U = beta * (V - Gamma * K S0),      A[j, l] = beta_j exp(G_j - G_l) (k_j . k_l)   (l < j)
        O = scale * (Gamma * Q S0 + tril(Q K^T * exp(G_i - G_j)) U)
        S_C = Gamma_C S0 + (K * exp(G_C - G))^T U

    All decay factors are ratios ``exp(G_i - G_j) <= 1`` (``g_log <= 0``), so nothing overflows.  Everything that does not
    depend on ``S0`` (A, its inverse applied to ``beta V`` and ``beta Gamma K``, the masked ``Q K^T``) is computed for all chunks
    at once; the sequential part is three small GEMMs per chunk."""
    B, T, H, K = q.shape
    HV, V = v.shape[2], v.shape[3]
    rep = HV // H
    C = chunk
    pad = (-T) % C
    f = lambda t: t.float()  # noqa: E731
    qf, kf = f(q).repeat_interleave(rep, 2), f(k).repeat_interleave(rep, 2)
    if l2norm:
        qf = qf * torch.rsqrt((qf * qf).sum(-1, keepdim=True) + 1e-6)
        kf = kf * torch.rsqrt((kf * kf).sum(-1, keepdim=True) + 1e-6)
    vf, gl, bt = f(v), f(g_log), f(beta)
    if pad:  # identity steps: no decay (g = 0), no write (beta = 0)
        z = lambda t: torch.cat([t, t.new_zeros(B, pad, *t.shape[2:])], 1)  # noqa: E731
        qf, kf, vf, gl, bt = z(qf), z(kf), z(vf), z(gl), z(bt)
    N = (T + pad) // C
    # [B, HV, N, C, *]
    r = lambda t: t.view(B, N, C, HV, -1).permute(0, 3, 1, 2, 4)  # noqa: E731
    Q, Kc, Vc = r(qf), r(kf), r(vf)
    G = r(gl[..., None])[..., 0].cumsum(-1)                      # cumulative log gate inside the chunk
    Bt = r(bt[..., None])[..., 0]
    Gam = torch.exp(G)
    ratio = torch.exp((G.unsqueeze(-1) - G.unsqueeze(-2)).clamp(max=0.0))  # exp(G_i - G_j), only i >= j is used
    KK = torch.matmul(Kc.to(compute_dtype), Kc.to(compute_dtype).transpose(-1, -2)).float()
    eye = torch.eye(C, device=q.device)
    A = (KK * ratio * Bt.unsqueeze(-1)).tril(-1)
    rhs = torch.cat([Vc * Bt.unsqueeze(-1), Kc * (Bt * Gam).unsqueeze(-1)], -1)      # [.., C, V + K]
    sol = torch.linalg.solve_triangular(eye + A, rhs, upper=False, unitriangular=True)
    UV, W = sol[..., :V], sol[..., V:]                                                # U = UV - W S0
    QK = (torch.matmul(Q.to(compute_dtype), Kc.to(compute_dtype).transpose(-1, -2)).float() * ratio).tril()
    Kdec = Kc * torch.exp(G[..., -1:] - G).unsqueeze(-1)                              # K * exp(G_C - G)
    QG = Q * Gam.unsqueeze(-1)
    out = torch.empty(B, HV, N, C, V, dtype=torch.float32, device=q.device)
    S = state.float().clone()                                                         # [B, HV, K, V]
    for c in range(N):
        U = UV[:, :, c] - torch.matmul(W[:, :, c], S)
        out[:, :, c] = torch.matmul(QG[:, :, c], S) + torch.matmul(QK[:, :, c], U)
        S = S * Gam[:, :, c, -1, None, None] + torch.matmul(Kdec[:, :, c].transpose(-1, -2), U)
    state.copy_(S.to(state.dtype))
    return (out * scale).permute(0, 2, 3, 1, 4).reshape(B, N * C, HV, V)[:, :T]


def _run(q, k, v, state, a, bgate, g_log, A_log, dt_bias, scale, l2norm, beta_is_logit, update_state, state_idx=None,
         cu_seqlens=None, out=None):
    B, T, H, K = q.shape
    HV, V = v.shape[2], v.shape[3]
    scale = scale if scale is not None else 1.0 / math.sqrt(K)
    if not q.is_cuda:
        if g_log is None:
            g_log = -torch.exp(A_log.float()) * torch.nn.functional.softplus(a.float() + dt_bias.float())
        beta = torch.sigmoid(bgate.float()) if beta_is_logit else bgate.float()
        if cu_seqlens is not None:
            res = torch.zeros(1, T, HV, V, dtype=torch.float32)
            cu = cu_seqlens.tolist()
            for i in range(len(cu) - 1):
                sl = slice(cu[i], cu[i + 1])
                st = state[i:i + 1] if update_state else state[i:i + 1].clone()
                res[:, sl] = gated_delta_rule_ref(q[:, sl], k[:, sl], v[:, sl], st, g_log[:, sl], beta[:, sl], scale, l2norm)
        else:
            idx = state_idx.long() if state_idx is not None else torch.arange(B)
            st = state[idx].clone()
            res = gated_delta_rule_ref(q, k, v, st, g_log, beta, scale, l2norm)
            if update_state:
                state[idx] = st
        res = res.to(q.dtype)
        if out is not None:
            out.copy_(res)
            return out
        return res
    q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
    if state.dtype != torch.float32 or not state.is_contiguous():
        raise ValueError("state must be a contiguous float32 tensor [N, HV, K, V]")
    res = out if out is not None else torch.empty(B, T, HV, V, dtype=q.dtype, device=q.device)
    nseq = (cu_seqlens.numel() - 1) if cu_seqlens is not None else B
    jit.load("ssm").call(
        "gated_delta_rule", state, q, k, v, a.to(q.dtype).contiguous() if a is not None else None,
        bgate.to(q.dtype).contiguous(), g_log.float().contiguous() if g_log is not None else None,
        A_log.float().contiguous() if A_log is not None else None, dt_bias.float().contiguous() if dt_bias is not None else None, res,
        state_idx.to(torch.int32).contiguous() if state_idx is not None else None,
        cu_seqlens.to(torch.int32).contiguous() if cu_seqlens is not None else None, float(scale), nseq, T, H, HV, K, V,
        1 if l2norm else 0, 1 if beta_is_logit else 0, 1 if update_state else 0, dtype_code(q.dtype), 1, stream_ptr(q))
    return res


def gated_delta_rule_decode(q, k, v, state, A_log, a, dt_bias, b, scale: Optional[float] = None,
                            output: Optional[torch.Tensor] = None, use_qk_l2norm: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """One decode step; ``state [B, HV, K, V]`` fp32 is updated in place.  Returns ``(output [B,1,HV,V], state)``."""
    o = _run(q, k, v, state, a, b, None, A_log, dt_bias, scale, use_qk_l2norm, True, True, out=output)
    return o, state


def gated_delta_rule_decode_pretranspose(q, k, v, state, A_log, a, dt_bias, b, scale: Optional[float] = None,
                                         output: Optional[torch.Tensor] = None, use_qk_l2norm: bool = True):
    """Same step for a V-major state ``[B, HV, V, K]`` (transposed on the fly; prefer the K-major entry point)."""
    st = state.transpose(-1, -2).contiguous()
    o = _run(q, k, v, st, a, b, None, A_log, dt_bias, scale, use_qk_l2norm, True, True, out=output)
    state.copy_(st.transpose(-1, -2))
    return o, state


def gated_delta_rule_mtp(q, k, v, initial_state, initial_state_indices, A_log, a, dt_bias, b, scale: Optional[float] = None,
                         output: Optional[torch.Tensor] = None, intermediate_states_buffer=None,
                         disable_state_update: Optional[bool] = None, use_qk_l2norm: bool = True):
    """T > 1 tokens per sequence (speculative verification); ``initial_state [pool, HV, K, V]`` indexed by
    ``initial_state_indices [B]``."""
    if intermediate_states_buffer is not None:
        raise NotImplementedError("intermediate state caching is not implemented")
    o = _run(q, k, v, initial_state, a, b, None, A_log, dt_bias, scale, use_qk_l2norm, True, not bool(disable_state_update),
             state_idx=initial_state_indices, out=output)
    return o, initial_state


def chunk_gated_delta_rule(q, k, v, g: Optional[torch.Tensor] = None, beta: Optional[torch.Tensor] = None,
                           scale: Optional[float] = None, initial_state: Optional[torch.Tensor] = None,
                           output_final_state: bool = False, cu_seqlens: Optional[torch.Tensor] = None,
                           use_qk_l2norm_in_kernel: bool = False, output: Optional[torch.Tensor] = None,
                           output_state: Optional[torch.Tensor] = None, state_checkpoints=None, checkpoint_cu_starts=None,
                           checkpoint_every_n_tokens: int = 0, chunked: Optional[bool] = None,
                           chunk_size: int = 64) -> Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
    """Prefill: ``q/k [total, H, K]``, ``v [total, HV, V]``, ``g`` (multiplicative forget gate) / ``beta`` ``[total, HV]``
    fp32, packed sequences described by ``cu_seqlens``.

    ``chunked=True`` (or ``FIB200_GDN_CHUNKED=1``) runs the chunk-parallel WY algorithm (:func:`gated_delta_rule_chunked`:
    batched GEMMs + a triangular solve per 64-token chunk) instead of the token-sequential kernel."""
    if state_checkpoints is not None:
        raise NotImplementedError("state checkpoints are not implemented")
    total, H, K = q.shape
    HV, V = v.shape[1], v.shape[2]
    dev = q.device
    if cu_seqlens is None:
        cu_seqlens = torch.tensor([0, total], dtype=torch.int32, device=dev)
    n = cu_seqlens.numel() - 1
    if H > HV:  # GQA on the value side: replicate v heads
        v = v.repeat_interleave(H // HV, 1)
        HV = H
    g_log = torch.log(g.float()) if g is not None else torch.zeros(total, HV, device=dev)
    bt = beta.float() if beta is not None else torch.ones(total, HV, device=dev)
    state = output_state if output_state is not None else torch.zeros(n, HV, K, V, dtype=torch.float32, device=dev)
    if initial_state is not None:
        state.copy_(initial_state)
    elif output_state is not None:
        state.zero_()
    import os

    if chunked is None:
        chunked = os.environ.get("FIB200_GDN_CHUNKED", "0") == "1"
    if chunked:
        sc = scale if scale is not None else 1.0 / math.sqrt(K)
        o = torch.empty(total, HV, V, dtype=torch.float32, device=dev)
        cu = cu_seqlens.tolist()
        for i in range(n):
            sl = slice(cu[i], cu[i + 1])
            if cu[i + 1] > cu[i]:
                o[sl] = gated_delta_rule_chunked(q[None, sl], k[None, sl], v[None, sl], state[i:i + 1], g_log[None, sl], bt[None, sl], sc,
                                                 use_qk_l2norm_in_kernel, chunk_size)[0]
        o = o.to(q.dtype)
        if output is not None:
            output.copy_(o)
            o = output
        return (o, state) if output_final_state else o
    o = _run(q[None], k[None], v[None], state, None, bt[None], g_log[None], None, None, scale, use_qk_l2norm_in_kernel, False, True,
             cu_seqlens=cu_seqlens, out=output[None] if output is not None else None)[0]
    return (o, state) if output_final_state else o
