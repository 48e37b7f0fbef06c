"""Mamba-2 decode engine: the state-space family on this library's ops (``mamba.selective_state_update`` for the recurrence, the
norm / activation / GEMM ops around it).

Per layer (one token per request; per-request rolling convolution state and SSM state addressed by a slot index):

    x     = rmsnorm(h)
    z, xBC, dt = in_proj . x                                   -> gate [H P], conv input [H P + 2 G N], step size [H]
    xBC   = silu(causal depthwise conv over the last ``d_conv`` inputs)      (conv state: shift in the new column)
    xs, B, C = split(xBC)
    y     = selective_state_update(state, xs, dt, A, B, C, D, z, dt_bias, softplus)     -> state <- exp(dt A) state + dt xs (x) B ;
                                                                                           y = (state . C + D xs) * silu(z)
    h    += out_proj . rmsnorm(y)                                             (the gated norm of Mamba-2, gate applied inside the update op)

Device agnostic like the other op-by-op engines; tested against a plain PyTorch recurrence."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch

from .. import norm
from ..gemm.dense import linear
from ..mamba import selective_state_update


@dataclass
class Mamba2Config:
    hidden_size: int = 2560
    num_layers: int = 64
    expand: int = 2
    head_dim: int = 64
    state_size: int = 128
    n_groups: int = 1
    conv_kernel: int = 4
    vocab_size: int = 50288
    rms_eps: float = 1e-5
    name: str = "mamba2-2.7b"

    @staticmethod
    def mamba2_2_7b() -> "Mamba2Config":
        return Mamba2Config()

    @staticmethod
    def tiny() -> "Mamba2Config":
        return Mamba2Config(hidden_size=128, num_layers=3, head_dim=32, state_size=16, vocab_size=256, name="mamba2-tiny")

    @property
    def d_inner(self) -> int:
        return self.expand * self.hidden_size

    @property
    def num_heads(self) -> int:
        return self.d_inner // self.head_dim

    @property
    def conv_dim(self) -> int:
        return self.d_inner + 2 * self.n_groups * self.state_size


class Mamba2DecodeEngine:
    """Random-init Mamba-2: ``plan(slots)`` binds the batch rows to state slots, ``step()`` consumes ``self.tokens``."""

    def __init__(self, cfg: Mamba2Config, max_slots: int, device: str = "cuda", dtype: torch.dtype = torch.bfloat16, seed: int = 0) -> None:
        self.cfg, self.device, self.dtype = cfg, torch.device(device), dtype
        g = torch.Generator(device="cpu").manual_seed(seed)
        w = lambda r, c: (torch.randn(r, c, generator=g) / c ** 0.5).to(dtype).to(self.device)  # noqa: E731
        ones = lambda n: (1.0 + 0.1 * torch.randn(n, generator=g)).to(dtype).to(self.device)  # noqa: E731
        h, hn, p, n = cfg.hidden_size, cfg.num_heads, cfg.head_dim, cfg.state_size
        self.embed = (torch.randn(cfg.vocab_size, h, generator=g) * 0.5).to(dtype).to(self.device)
        self.lm_head, self.final_norm = w(cfg.vocab_size, h), ones(h)
        self.layers: List[dict] = []
        for _ in range(cfg.num_layers):
            a = -torch.exp(torch.randn(hn, generator=g) * 0.5)                                       # one negative decay rate per head
            self.layers.append({
                "ln": ones(h), "in_proj": w(2 * cfg.d_inner + 2 * cfg.n_groups * n + hn, h),
                "conv_w": (torch.randn(cfg.conv_dim, cfg.conv_kernel, generator=g) * 0.3).to(dtype).to(self.device),
                "conv_b": (torch.randn(cfg.conv_dim, generator=g) * 0.1).to(dtype).to(self.device),
                "A": a[:, None, None].expand(hn, p, n).contiguous().float().to(self.device),
                "D": torch.randn(hn, generator=g)[:, None].expand(hn, p).contiguous().float().to(self.device),
                "dt_bias": (torch.randn(hn, generator=g) * 0.5)[:, None].expand(hn, p).contiguous().float().to(self.device),
                "gate_norm": ones(cfg.d_inner), "out_proj": w(h, cfg.d_inner),
                "conv_state": torch.zeros(max_slots, cfg.conv_dim, cfg.conv_kernel, dtype=dtype, device=self.device),
                "ssm_state": torch.zeros(max_slots, hn, p, n, dtype=torch.float32, device=self.device)})
        self.logits: Optional[torch.Tensor] = None

    def plan(self, slots: torch.Tensor) -> None:
        self.slots = slots.to(self.device, torch.int32)
        self.batch = b = self.slots.numel()
        self.tokens = torch.zeros(b, dtype=torch.int64, device=self.device)
        self.next_tokens = torch.zeros(b, dtype=torch.int64, device=self.device)

    def _mixer(self, l: dict, x: torch.Tensor) -> torch.Tensor:
        cfg, b = self.cfg, self.batch
        hn, p, n, g = cfg.num_heads, cfg.head_dim, cfg.state_size, cfg.n_groups
        proj = linear(x, l["in_proj"])
        z, xbc, dt = proj[:, : cfg.d_inner], proj[:, cfg.d_inner: cfg.d_inner + cfg.conv_dim], proj[:, cfg.d_inner + cfg.conv_dim:]
        idx = self.slots.long()
        window = torch.cat([l["conv_state"][idx][:, :, 1:], xbc.unsqueeze(-1)], -1)                  # shift the new column in
        l["conv_state"][idx] = window
        xbc = torch.nn.functional.silu((window.float() * l["conv_w"].float()).sum(-1) + l["conv_b"].float()).to(self.dtype)
        xs = xbc[:, : cfg.d_inner].reshape(b, hn, p)
        bm = xbc[:, cfg.d_inner: cfg.d_inner + g * n].reshape(b, g, n)
        cm = xbc[:, cfg.d_inner + g * n:].reshape(b, g, n)
        y = selective_state_update(l["ssm_state"], xs.contiguous(), dt[:, :, None].expand(b, hn, p).contiguous(), l["A"], bm.contiguous(),
                                   cm.contiguous(), l["D"], z.reshape(b, hn, p).contiguous(), l["dt_bias"], True, state_batch_indices=self.slots)
        return linear(norm.rmsnorm(y.reshape(b, cfg.d_inner).to(self.dtype), l["gate_norm"], cfg.rms_eps), l["out_proj"])

    def step(self) -> torch.Tensor:
        cfg = self.cfg
        res = self.embed[self.tokens]
        for l in self.layers:
            res = res + self._mixer(l, norm.rmsnorm(res, l["ln"], cfg.rms_eps))
        self.logits = linear(norm.rmsnorm(res, self.final_norm, cfg.rms_eps), self.lm_head)
        torch.argmax(self.logits, dim=-1, out=self.next_tokens)
        return self.next_tokens
