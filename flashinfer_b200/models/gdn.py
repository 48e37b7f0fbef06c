"""Gated-delta-net (linear attention) decode engine - the recurrent layer of Qwen3-Next style hybrids on ``gdn.gated_delta_rule_mtp``.

Per layer (one token per request; rolling conv state and a [HV, K, V] fp32 delta-rule state per slot):

    x = rmsnorm(h);   q, k, v, z, b, a = in_proj . x
    q, k, v = silu(causal depthwise conv over the last ``conv_kernel`` inputs)
    o = gated_delta_rule(q, k, v; state)      g = exp(-exp(A_log) softplus(a + dt_bias)),  beta = sigmoid(b),  q / k l2-normalised:
                                              S <- g S ;  S <- S + k (beta (v - k^T S))^T ;  o = (q / sqrt(K))^T S
    h += out_proj . (rmsnorm_per_head(o) * silu(z));   h += SwiGLU MLP(rmsnorm(h))

Full-attention layers of a hybrid stack are what ``models.transformer`` provides; this engine is the recurrent part alone."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch

from .. import activation, norm
from ..gdn import gated_delta_rule_mtp
from ..gemm.dense import linear


@dataclass
class GDNConfig:
    hidden_size: int = 2048
    num_layers: int = 36
    num_k_heads: int = 16
    num_v_heads: int = 32
    head_k_dim: int = 128
    head_v_dim: int = 128
    conv_kernel: int = 4
    intermediate_size: int = 5632
    vocab_size: int = 151936
    rms_eps: float = 1e-6
    name: str = "gdn"

    @staticmethod
    def tiny() -> "GDNConfig":
        return GDNConfig(hidden_size=128, num_layers=2, num_k_heads=2, num_v_heads=4, head_k_dim=16, head_v_dim=8, intermediate_size=192,
                         vocab_size=256, name="gdn-tiny")

    @property
    def qkv_dim(self) -> int:
        return 2 * self.num_k_heads * self.head_k_dim + self.num_v_heads * self.head_v_dim


class GDNDecodeEngine:
    def __init__(self, cfg: GDNConfig, max_slots: int, device: str = "cuda", dtype: torch.dtype = torch.bfloat16, seed: int = 0) -> None:
        self.cfg, self.device, self.dtype = cfg, torch.device(device), dtype
        g = torch.Generator(device="cpu").manual_seed(seed)
        w = lambda r, c: (torch.randn(r, c, generator=g) / c ** 0.5).to(dtype).to(self.device)  # noqa: E731
        ones = lambda n: (1.0 + 0.1 * torch.randn(n, generator=g)).to(dtype).to(self.device)  # noqa: E731
        h, hv = cfg.hidden_size, cfg.num_v_heads
        vdim = hv * cfg.head_v_dim
        self.embed = (torch.randn(cfg.vocab_size, h, generator=g) * 0.5).to(dtype).to(self.device)
        self.lm_head, self.final_norm = w(cfg.vocab_size, h), ones(h)
        self.layers: List[dict] = []
        for _ in range(cfg.num_layers):
            self.layers.append({
                "ln1": ones(h), "ln2": ones(h), "in_proj": w(cfg.qkv_dim + vdim + 2 * hv, h),
                "conv_w": (torch.randn(cfg.qkv_dim, cfg.conv_kernel, generator=g) * 0.3).to(dtype).to(self.device),
                "A_log": (torch.randn(hv, generator=g) * 0.5).float().to(self.device), "dt_bias": (torch.randn(hv, generator=g) * 0.1).float().to(self.device),
                "o_norm": ones(cfg.head_v_dim), "out_proj": w(h, vdim), "w_gu": w(2 * cfg.intermediate_size, h), "w_d": w(h, cfg.intermediate_size),
                "conv_state": torch.zeros(max_slots, cfg.qkv_dim, cfg.conv_kernel, dtype=dtype, device=self.device),
                "state": torch.zeros(max_slots, hv, cfg.head_k_dim, cfg.head_v_dim, dtype=torch.float32, device=self.device)})
        self.logits: Optional[torch.Tensor] = None

    def plan(self, slots: torch.Tensor) -> None:
        self.slots = slots.to(self.device, torch.int32)
        self.batch = b = self.slots.numel()
        self.tokens = torch.zeros(b, dtype=torch.int64, device=self.device)
        self.next_tokens = torch.zeros(b, dtype=torch.int64, device=self.device)

    def _mixer(self, l: dict, x: torch.Tensor) -> torch.Tensor:
        cfg, b = self.cfg, self.batch
        hk, hv, kd, vd = cfg.num_k_heads, cfg.num_v_heads, cfg.head_k_dim, cfg.head_v_dim
        proj = linear(x, l["in_proj"])
        qkv, z = proj[:, : cfg.qkv_dim], proj[:, cfg.qkv_dim: cfg.qkv_dim + hv * vd]
        bgate, a = proj[:, cfg.qkv_dim + hv * vd: cfg.qkv_dim + hv * vd + hv], proj[:, cfg.qkv_dim + hv * vd + hv:]
        idx = self.slots.long()
        window = torch.cat([l["conv_state"][idx][:, :, 1:], qkv.unsqueeze(-1)], -1)
        l["conv_state"][idx] = window
        qkv = torch.nn.functional.silu((window.float() * l["conv_w"].float()).sum(-1)).to(self.dtype)
        q = qkv[:, : hk * kd].reshape(b, 1, hk, kd)
        k = qkv[:, hk * kd: 2 * hk * kd].reshape(b, 1, hk, kd)
        v = qkv[:, 2 * hk * kd:].reshape(b, 1, hv, vd)
        o, _ = gated_delta_rule_mtp(q.contiguous(), k.contiguous(), v.contiguous(), l["state"], self.slots, l["A_log"], a.reshape(b, 1, hv).contiguous(),
                                    l["dt_bias"], bgate.reshape(b, 1, hv).contiguous(), disable_state_update=False, state_layout="KV")
        o = norm.rmsnorm(o.reshape(b * hv, vd).to(self.dtype), l["o_norm"], cfg.rms_eps).view(b, hv * vd)
        gated = (o.float() * torch.nn.functional.silu(z.float())).to(self.dtype)
        return linear(gated, l["out_proj"])

    def step(self) -> torch.Tensor:
        cfg = self.cfg
        res = self.embed[self.tokens]
        for l in self.layers:
            res = res + self._mixer(l, norm.rmsnorm(res, l["ln1"], cfg.rms_eps))
            res = res + linear(activation.silu_and_mul(linear(norm.rmsnorm(res, l["ln2"], cfg.rms_eps), l["w_gu"])), l["w_d"])
        self.logits = linear(norm.rmsnorm(res, self.final_norm, cfg.rms_eps), self.lm_head)
        torch.argmax(self.logits, dim=-1, out=self.next_tokens)
        return self.next_tokens
