"""Plain-PyTorch fp32 oracles for every op.

These serve two purposes:
* the CPU execution path of the public API (``single_prefill_with_kv_cache`` on CPU tensors is
  BASELINE.json config #1 — the reference itself has no CPU path);
* the numerics oracle of every GPU test (tests compare the CUDA kernel against these).

Conventions follow the reference: LSE in base-2 (include/flashinfer/attention/state.cuh:46),
``window_left=-1`` = no window, ``logits_soft_cap=0`` = off.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

LOG2E = 1.4426950408889634


# ------------------------------------------------------------------ attention
def attention_ref(
    q: torch.Tensor,  # [qo_len, Hq, D]
    k: torch.Tensor,  # [kv_len, Hkv, D]
    v: torch.Tensor,  # [kv_len, Hkv, Dv]
    causal: bool = False,
    sm_scale: Optional[float] = None,
    logits_soft_cap: float = 0.0,
    window_left: int = -1,
    custom_mask: Optional[torch.Tensor] = None,  # [qo_len, kv_len] bool
    alibi_slopes: Optional[torch.Tensor] = None,  # [Hq]
    sinks: Optional[torch.Tensor] = None,  # [Hq] attention-sink logits
    return_lse: bool = True,
) -> Tuple[torch.Tensor, torch.Tensor]:
    qo_len, hq, d = q.shape
    kv_len, hkv, _ = k.shape
    dv = v.shape[-1]
    group = hq // hkv
    if sm_scale is None:
        sm_scale = 1.0 / math.sqrt(d)
    qf = q.float()
    kf = k.float().repeat_interleave(group, dim=1)
    vf = v.float().repeat_interleave(group, dim=1)
    logits = torch.einsum("qhd,khd->hqk", qf, kf) * sm_scale
    if logits_soft_cap > 0:
        logits = logits_soft_cap * torch.tanh(logits / logits_soft_cap)
    qpos = torch.arange(qo_len, device=q.device)[:, None] + (kv_len - qo_len)
    kpos = torch.arange(kv_len, device=q.device)[None, :]
    if alibi_slopes is not None:
        logits = logits + alibi_slopes.float()[:, None, None] * (kpos - qpos)[None].float()
    mask = torch.ones(qo_len, kv_len, dtype=torch.bool, device=q.device)
    if causal:
        mask &= kpos <= qpos
    if window_left >= 0:
        mask &= kpos >= qpos - window_left
    if custom_mask is not None:
        mask &= custom_mask.to(torch.bool).reshape(qo_len, kv_len)
    logits = logits.masked_fill(~mask[None], float("-inf"))
    if sinks is not None:
        sink = sinks.float()[:, None, None].expand(hq, qo_len, 1)
        full = torch.cat([logits, sink], dim=-1)
        lse_e = torch.logsumexp(full, dim=-1)
    else:
        lse_e = torch.logsumexp(logits, dim=-1)  # [hq, qo]
    p = torch.exp(logits - torch.where(torch.isinf(lse_e), torch.zeros_like(lse_e), lse_e)[..., None])
    p = torch.where(torch.isinf(lse_e)[..., None], torch.zeros_like(p), p)
    out = torch.einsum("hqk,khd->qhd", p, vf)
    lse = (lse_e * LOG2E).transpose(0, 1).contiguous()  # [qo, hq], base 2
    return out.to(q.dtype) if q.dtype != torch.float32 else out, lse


def gather_paged_kv(
    k_cache: torch.Tensor,
    v_cache: torch.Tensor,
    kv_indices: torch.Tensor,
    kv_indptr: torch.Tensor,
    kv_last_page_len: torch.Tensor,
    req: int,
    kv_layout: str = "NHD",
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Materialise request `req`'s contiguous K/V [kv_len, H, D] from the paged cache."""
    s, e = int(kv_indptr[req]), int(kv_indptr[req + 1])
    if e == s:
        h = k_cache.shape[2] if kv_layout == "NHD" else k_cache.shape[1]
        return (k_cache.new_zeros(0, h, k_cache.shape[-1]), v_cache.new_zeros(0, h, v_cache.shape[-1]))
    pages = kv_indices[s:e].long()
    kp, vp = k_cache[pages], v_cache[pages]
    if kv_layout == "HND":
        kp, vp = kp.transpose(1, 2), vp.transpose(1, 2)
    page_size = kp.shape[1]
    kv_len = (e - s - 1) * page_size + int(kv_last_page_len[req])
    k = kp.reshape(-1, kp.shape[2], kp.shape[3])[:kv_len]
    v = vp.reshape(-1, vp.shape[2], vp.shape[3])[:kv_len]
    return k, v


def batch_paged_attention_ref(
    q: torch.Tensor,  # [total_q, Hq, D]
    qo_indptr: torch.Tensor,
    k_cache,
    v_cache,
    kv_indptr,
    kv_indices,
    kv_last_page_len,
    kv_layout="NHD",
    causal=True,
    sm_scale=None,
    logits_soft_cap=0.0,
    window_left=-1,
):
    outs, lses = [], []
    batch = qo_indptr.numel() - 1
    for b in range(batch):
        qs, qe = int(qo_indptr[b]), int(qo_indptr[b + 1])
        k, v = gather_paged_kv(k_cache, v_cache, kv_indices, kv_indptr, kv_last_page_len, b, kv_layout)
        if qe == qs:
            continue
        if k.shape[0] == 0:
            outs.append(torch.zeros(qe - qs, q.shape[1], v_cache.shape[-1], dtype=q.dtype, device=q.device))
            lses.append(torch.full((qe - qs, q.shape[1]), float("-inf"), device=q.device))
            continue
        o, l = attention_ref(q[qs:qe], k, v, causal, sm_scale, logits_soft_cap, window_left)
        outs.append(o)
        lses.append(l)
    if not outs:                                           # an empty batch
        return (torch.zeros(0, q.shape[1], v_cache.shape[-1], dtype=q.dtype, device=q.device),
                torch.zeros(0, q.shape[1], dtype=torch.float32, device=q.device))
    return torch.cat(outs, 0), torch.cat(lses, 0)


def merge_state_ref(v_a, s_a, v_b, s_b):
    """(o, lse base-2) merge operator (reference include/flashinfer/attention/state.cuh:31-78)."""
    sa, sb = s_a.float(), s_b.float()
    m = torch.maximum(sa, sb)
    m = torch.where(torch.isinf(m), torch.zeros_like(m), m)
    wa, wb = torch.exp2(sa - m), torch.exp2(sb - m)
    den = wa + wb
    o = (v_a.float() * wa[..., None] + v_b.float() * wb[..., None]) / torch.where(den > 0, den, torch.ones_like(den))[..., None]
    s = torch.log2(den) + m
    return o.to(v_a.dtype), s


def merge_states_ref(v, s):
    """v: [n, num_states, H, D], s: [n, num_states, H] -> ([n,H,D], [n,H])."""
    sf = s.float()
    m = sf.max(dim=1, keepdim=True).values
    m = torch.where(torch.isinf(m), torch.zeros_like(m), m)
    w = torch.exp2(sf - m)
    den = w.sum(1)
    o = (v.float() * w[..., None]).sum(1) / torch.where(den > 0, den, torch.ones_like(den))[..., None]
    return o.to(v.dtype), torch.log2(den) + m.squeeze(1)


# ------------------------------------------------------------------ norm / activation
def rmsnorm_ref(x, w, eps=1e-6, weight_bias: float = 0.0):
    xf = x.float()
    var = xf.pow(2).mean(-1, keepdim=True)
    return (xf * torch.rsqrt(var + eps) * (w.float() + weight_bias)).to(x.dtype)


def fused_add_rmsnorm_ref(x, residual, w, eps=1e-6, weight_bias: float = 0.0):
    r = (x.float() + residual.float()).to(x.dtype)
    return rmsnorm_ref(r, w, eps, weight_bias), r


def layernorm_ref(x, gamma, beta, eps=1e-6):
    xf = x.float()
    mu = xf.mean(-1, keepdim=True)
    var = (xf - mu).pow(2).mean(-1, keepdim=True)
    return ((xf - mu) * torch.rsqrt(var + eps) * gamma.float() + beta.float()).to(x.dtype)


def silu_and_mul_ref(x):
    d = x.shape[-1] // 2
    return (torch.nn.functional.silu(x[..., :d].float()) * x[..., d:].float()).to(x.dtype)


def gelu_and_mul_ref(x, approximate="none"):
    d = x.shape[-1] // 2
    return (torch.nn.functional.gelu(x[..., :d].float(), approximate=approximate) * x[..., d:].float()).to(x.dtype)


# ------------------------------------------------------------------ rope
def rope_freqs(rotary_dim, rope_scale=1.0, rope_theta=1e4, device="cpu", llama31=None):
    inv = 1.0 / (rope_theta ** (torch.arange(0, rotary_dim, 2, device=device, dtype=torch.float32) / rotary_dim))
    if llama31 is not None:
        low_freq_factor, high_freq_factor, old_context_len = llama31
        smooth_a = old_context_len / (2 * math.pi * (high_freq_factor - low_freq_factor))
        smooth_b = -1.0 / (high_freq_factor / low_freq_factor - 1.0)
        smooth = torch.clamp(inv * smooth_a + smooth_b, 0.0, 1.0)
        inv = (1 - smooth) * (inv / rope_scale) + smooth * inv
    else:
        inv = inv / rope_scale
    return inv


def apply_rope_ref(x, pos, rotary_dim=None, interleave=False, rope_scale=1.0, rope_theta=1e4, llama31=None):
    """x: [n, H, D], pos: [n] -> rotated x (fp32 math)."""
    d = x.shape[-1]
    rd = rotary_dim or d
    inv = rope_freqs(rd, rope_scale, rope_theta, x.device, llama31)
    ang = pos.float()[:, None] * inv[None, :]  # [n, rd/2]
    cos, sin = torch.cos(ang)[:, None, :], torch.sin(ang)[:, None, :]
    xf = x.float()
    xr = xf[..., :rd]
    if interleave:
        x1, x2 = xr[..., 0::2], xr[..., 1::2]
        o1, o2 = x1 * cos - x2 * sin, x2 * cos + x1 * sin
        rot = torch.stack([o1, o2], dim=-1).flatten(-2)
    else:
        x1, x2 = xr[..., : rd // 2], xr[..., rd // 2 :]
        rot = torch.cat([x1 * cos - x2 * sin, x2 * cos + x1 * sin], dim=-1)
    out = torch.cat([rot, xf[..., rd:]], dim=-1)
    return out.to(x.dtype)


def apply_rope_cos_sin_cache_ref(x, pos, cos_sin_cache, interleave=False):
    """cos_sin_cache [max_pos, rotary_dim] = [cos | sin] halves."""
    rd = cos_sin_cache.shape[-1]
    cs = cos_sin_cache[pos.long()].float()
    cos, sin = cs[:, None, : rd // 2], cs[:, None, rd // 2 :]
    xf = x.float()
    xr = xf[..., :rd]
    if interleave:
        x1, x2 = xr[..., 0::2], xr[..., 1::2]
        rot = torch.stack([x1 * cos - x2 * sin, x2 * cos + x1 * sin], dim=-1).flatten(-2)
    else:
        x1, x2 = xr[..., : rd // 2], xr[..., rd // 2 :]
        rot = torch.cat([x1 * cos - x2 * sin, x2 * cos + x1 * sin], dim=-1)
    return torch.cat([rot, xf[..., rd:]], dim=-1).to(x.dtype)
