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


def _check_layout(state_layout: str) -> bool:
    if state_layout not in ("VK", "KV"):
        raise ValueError(f"state_layout must be 'VK' (reference: K-last) or 'KV' (native: K-major), got {state_layout!r}")
    return state_layout == "VK"


def gated_delta_rule_decode(q, k, v, state, A_log, a, dt_bias, b, scale: Optional[float] = None,
                            output: Optional[torch.Tensor] = None, use_qk_l2norm: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """One decode step; ``state [B, HV, K, V]`` fp32 (K-major, the kernel's native layout) is updated in place.  Returns
    ``(output [B,1,HV,V], state)``."""
    o = _run(q, k, v, state, a, b, None, A_log, dt_bias, scale, use_qk_l2norm, True, True, out=output)
    return o, state


def gated_delta_rule_decode_pretranspose(q, k, v, state, A_log, a, dt_bias, b, scale: Optional[float] = None,
                                         output: Optional[torch.Tensor] = None, use_qk_l2norm: bool = True,
                                         initial_state: Optional[torch.Tensor] = None, initial_state_indices: Optional[torch.Tensor] = None,
                                         output_state_indices: Optional[torch.Tensor] = None):
    """Same step for a V-major (K-last) state ``[B, HV, V, K]``, transposed on the fly (prefer the K-major entry point).

    Pool form (reference gdn_decode.py :118): ``state=None``, ``initial_state [pool, HV, V, K]`` gathered through
    ``initial_state_indices [B]`` and written back to ``output_state_indices`` (default: the same slots).  Entries whose index is
    ``-1`` are padding: their slot is not touched and their output row is zero (the reference's float32-path semantics)."""
    if initial_state is None:
        if state is None:
            raise ValueError("gated_delta_rule_decode_pretranspose: pass state, or initial_state with initial_state_indices")
        if initial_state_indices is not None or output_state_indices is not None:
            raise ValueError("initial_state_indices / output_state_indices need the initial_state pool")
        st = state.float().transpose(-1, -2).contiguous()
        o = _run(q, k, v, st, a, b, None, A_log, dt_bias, scale, use_qk_l2norm, True, True, out=output)
        state.copy_(st.transpose(-1, -2))
        return o, state
    if state is not None:
        raise ValueError("gated_delta_rule_decode_pretranspose: state and initial_state are mutually exclusive")
    if initial_state_indices is None:
        raise ValueError("initial_state needs initial_state_indices")
    src = initial_state_indices.long()
    dst = output_state_indices.long() if output_state_indices is not None else src
    live = (src >= 0) & (dst >= 0)
    st = initial_state[src.clamp(min=0)].float().transpose(-1, -2).contiguous()
    o = _run(q, k, v, st, a, b, None, A_log, dt_bias, scale, use_qk_l2norm, True, True, out=output)
    o.mul_(live.view(-1, 1, 1, 1).to(o.dtype))
    initial_state.index_copy_(0, dst[live], st.transpose(-1, -2)[live].to(initial_state.dtype))
    return o, initial_state


def gated_delta_rule_mtp(q, k, v, initial_state, initial_state_indices, A_log, a, dt_bias, b, scale: Optional[float] = None,
                         output: Optional[torch.Tensor] = None, intermediate_states_buffer: Optional[torch.Tensor] = None,
                         disable_state_update: Optional[bool] = None, use_qk_l2norm: bool = True, state_layout: str = "VK"):
    """T > 1 tokens per sequence (speculative verification); the state pool is indexed by ``initial_state_indices [B]``.

    ``state_layout="VK"`` (default, the reference's convention): ``initial_state [pool, HV, V, K]`` (K-last); the touched slots are
    gathered and transposed around the kernel.  ``state_layout="KV"``: ``[pool, HV, K, V]``, the kernel's native layout - it
    indexes the pool itself, nothing is copied (what ``models.gdn`` uses).  ``disable_state_update=None`` means True with a
    warning, like in the reference (its default flips in a later release: pass it explicitly).  ``intermediate_states_buffer
    [>= B, >= T, HV, V, K]`` receives the state after every token (row ``b`` of the call, not the pool slot), for roll-back after
    partial acceptance; with it the tokens run as T single-token launches."""
    vk = _check_layout(state_layout)
    if disable_state_update is None:
        import warnings

        warnings.warn("gated_delta_rule_mtp(): disable_state_update defaults to True (the state pool is NOT updated); pass it "
                      "explicitly - the reference flips this default in 0.7.0", FutureWarning, stacklevel=2)
        disable_state_update = True
    update = not disable_state_update
    B, T = q.shape[0], q.shape[1]
    if intermediate_states_buffer is None and not vk:
        o = _run(q, k, v, initial_state, a, b, None, A_log, dt_bias, scale, use_qk_l2norm, True, update,
                 state_idx=initial_state_indices, out=output)
        return o, initial_state
    idx = initial_state_indices.long()
    st = initial_state[idx].float()
    st = st.transpose(-1, -2).contiguous() if vk else st.contiguous()           # working copy [B, HV, K, V]
    if intermediate_states_buffer is None:
        o = _run(q, k, v, st, a, b, None, A_log, dt_bias, scale, use_qk_l2norm, True, True, out=output)
    else:
        buf = intermediate_states_buffer
        if buf.shape[0] < B or buf.shape[1] < T:
            raise ValueError(f"intermediate_states_buffer {tuple(buf.shape)} must hold at least [B={B}, T={T}] states")
        o = output if output is not None else torch.empty(B, T, v.shape[2], v.shape[3], dtype=q.dtype, device=q.device)
        for t in range(T):
            sl = slice(t, t + 1)
            o[:, sl] = _run(q[:, sl], k[:, sl], v[:, sl], st, a[:, sl], b[:, sl], None, A_log, dt_bias, scale, use_qk_l2norm, True, True)
            buf[:B, t] = st.transpose(-1, -2).to(buf.dtype)
    if update:
        initial_state.index_copy_(0, idx, (st.transpose(-1, -2) if vk else st).to(initial_state.dtype))
    return o, initial_state


def chunk_gated_delta_rule(q, k, v, g: Optional[torch.Tensor] = None, beta: Optional[torch.Tensor] = None,
                           scale: Optional[float] = None, initial_state: Optional[torch.Tensor] = None,
                           output_final_state: bool = False, cu_seqlens: Optional[torch.Tensor] = None,
                           use_qk_l2norm_in_kernel: bool = False, output: Optional[torch.Tensor] = None,
                           output_state: Optional[torch.Tensor] = None, state_checkpoints: Optional[torch.Tensor] = None,
                           checkpoint_cu_starts: Optional[torch.Tensor] = None, checkpoint_every_n_tokens: int = 0,
                           chunked: Optional[bool] = None, chunk_size: int = 64,
                           state_layout: str = "VK") -> Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
    """Prefill: ``q/k [total, H, K]``, ``v [total, HV, V]``, ``g`` (multiplicative forget gate) / ``beta`` ``[total, HV]``
    fp32, packed sequences described by ``cu_seqlens``.

    States (``initial_state``, ``output_state`` / the returned final state, ``state_checkpoints``) are ``[*, HV, V, K]`` (K-last,
    the reference's layout: gdn_prefill.py :186) by default and transposed around the kernel; ``state_layout="KV"`` takes and
    returns the kernel's native ``[*, HV, K, V]`` without copies.

    ``checkpoint_every_n_tokens = n > 0`` (a multiple of 64): the state after every n tokens of sequence ``i`` is written to
    ``state_checkpoints[checkpoint_cu_starts[i] + j]`` (``seq_len_i // n`` checkpoints per sequence).  The sequences then run in
    rounds of n tokens - round j processes the j-th segment of every sequence that has one, as one packed launch.

    ``chunked=True`` (or ``FIB200_GDN_CHUNKED=1``) runs the chunk-parallel WY algorithm (:func:`gated_delta_rule_chunked`:
    batched GEMMs + a triangular solve per 64-token chunk) instead of the token-sequential kernel."""
    vk = _check_layout(state_layout)
    if checkpoint_every_n_tokens < 0:
        raise ValueError(f"checkpoint_every_n_tokens must be non-negative, got {checkpoint_every_n_tokens}")
    if checkpoint_every_n_tokens > 0:
        if checkpoint_every_n_tokens % 64 != 0:
            raise ValueError(f"checkpoint_every_n_tokens must be a multiple of the chunk size (64), got {checkpoint_every_n_tokens}")
        if state_checkpoints is None or checkpoint_cu_starts is None:
            raise ValueError("state_checkpoints and checkpoint_cu_starts must both be provided when checkpoint_every_n_tokens > 0")
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
    native_out = output_state is not None and not vk                      # the kernel may work in the caller's buffer
    state = output_state if native_out else torch.zeros(n, HV, K, V, dtype=torch.float32, device=dev)
    if initial_state is not None:
        state.copy_(initial_state.transpose(-1, -2) if vk else initial_state)
    elif native_out:
        state.zero_()
    import os

    if chunked is None:
        chunked = os.environ.get("FIB200_GDN_CHUNKED", "0") == "1"

    def segment(qs, ks, vs, gs, bs, st, cu, out):
        """One packed launch: tokens ``[0, cu[-1])`` of the given tensors, sequence ``i`` continuing ``st[i]`` (updated in place)."""
        if chunked:
            sc = scale if scale is not None else 1.0 / math.sqrt(K)
            o = torch.empty(qs.shape[0], HV, V, dtype=torch.float32, device=dev)
            cl = cu.tolist()
            for i in range(len(cl) - 1):
                sl = slice(cl[i], cl[i + 1])
                if cl[i + 1] > cl[i]:
                    o[sl] = gated_delta_rule_chunked(qs[None, sl], ks[None, sl], vs[None, sl], st[i:i + 1], gs[None, sl], bs[None, sl], sc,
                                                     use_qk_l2norm_in_kernel, chunk_size)[0]
            o = o.to(q.dtype)
            if out is not None:
                out.copy_(o)
                o = out
            return o
        return _run(qs[None], ks[None], vs[None], st, None, bs[None], gs[None], None, None, scale, use_qk_l2norm_in_kernel, False, True,
                    cu_seqlens=cu, out=out[None] if out is not None else None)[0]

    if checkpoint_every_n_tokens == 0:
        o = segment(q, k, v, g_log, bt, state, cu_seqlens, output)
    else:
        step = checkpoint_every_n_tokens
        cu = cu_seqlens.tolist()
        starts = checkpoint_cu_starts.tolist()
        lens = [cu[i + 1] - cu[i] for i in range(n)]
        for i in range(n):
            if starts[i + 1] - starts[i] != lens[i] // step:
                raise ValueError(f"checkpoint_cu_starts: sequence {i} of {lens[i]} tokens has {lens[i] // step} checkpoints, "
                                 f"got {starts[i + 1] - starts[i]}")
        o = output if output is not None else torch.empty(total, HV, V, dtype=q.dtype, device=dev)
        for j in range((max(lens) + step - 1) // step if lens else 0):
            act = [i for i in range(n) if lens[i] > j * step]
            seg = [min(step, lens[i] - j * step) for i in act]
            tok = torch.cat([torch.arange(cu[i] + j * step, cu[i] + j * step + m, device=dev) for i, m in zip(act, seg)])
            cu_j = torch.tensor([0] + torch.tensor(seg).cumsum(0).tolist(), dtype=torch.int32, device=dev)
            act_t = torch.tensor(act, device=dev)
            st = state[act_t].contiguous()
            o[tok] = segment(q[tok], k[tok], v[tok], g_log[tok], bt[tok], st, cu_j, None).to(o.dtype)
            state[act_t] = st
            full = [x for x, (i, m) in enumerate(zip(act, seg)) if m == step]
            if full:
                rows = torch.tensor([starts[act[x]] + j for x in full], device=dev)
                snap = st[torch.tensor(full, device=dev)]
                state_checkpoints[rows] = (snap.transpose(-1, -2) if vk else snap).to(state_checkpoints.dtype)
    if native_out:
        return (o, state) if output_final_state else o
    final = state.transpose(-1, -2) if vk else state
    if output_state is not None:
        output_state.copy_(final)
        final = output_state
    return (o, final.contiguous()) if output_final_state else o
