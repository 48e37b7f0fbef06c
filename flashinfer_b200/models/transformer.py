"""A configurable GQA decoder covering the mainstream open model families with one op-by-op decode engine built from the public ops:

* **Mixtral / Qwen-MoE**: routed SwiGLU experts, top-k of the raw router logits + softmax over the selected ("renormalize");
* **Qwen3**: per-head RMSNorm of q and k before RoPE;
* **Gemma-2 / 3**: ``(1 + w)`` RMSNorm weights, post-attention / post-FFN norms, GeGLU (tanh), embedding scale ``sqrt(hidden)``,
  attention logit soft-cap, alternating sliding-window / global layers, final-logit soft-cap, custom query scale;
* **Llama / Mistral**: the plain configuration (the fused flagship path for Llama is ``models.llama``).

Per layer: RMSNorm -> QKV GEMM -> [q/k norm] -> RoPE -> paged KV append -> paged decode attention (global or sliding-window plan, with
soft-cap) -> O GEMM -> [post norm] -> residual -> RMSNorm -> dense gated MLP or MoE -> [post norm] -> residual.  Device agnostic like
``models.deepseek`` (native kernels on CUDA, eager paths on CPU; tested against a plain PyTorch model)."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import torch

from .. import activation, norm, page, rope
from ..decode import BatchDecodeWithPagedKVCacheWrapper
from ..prefill import BatchPrefillWithPagedKVCacheWrapper
from ..fused_moe.core import RoutingMethodType, moe_forward, route
from ..gemm.dense import linear
from ..parallel import all_reduce_fp32, shard_cols, shard_rows


@dataclass
class TransformerConfig:
    hidden_size: int = 4096
    intermediate_size: int = 14336
    num_layers: int = 32
    num_qo_heads: int = 32
    num_kv_heads: int = 8
    head_dim: int = 128
    vocab_size: int = 32000
    rms_eps: float = 1e-5
    rope_theta: float = 1e6
    activation: str = "silu"                      # "silu" (SwiGLU) or "gelu_tanh" (GeGLU)
    num_experts: int = 0                          # 0 = dense MLP
    num_experts_per_tok: int = 2
    qk_norm: bool = False                         # Qwen3: RMSNorm over head_dim on q and k
    gemma_norm: bool = False                      # weights stored as (w - 1)
    post_norms: bool = False                      # Gemma-2: norms after attention and after the MLP
    embed_scale: float = 1.0
    query_pre_attn_scalar: Optional[float] = None  # softmax scale = this ** -0.5 (default head_dim)
    attn_logit_softcap: float = 0.0
    final_logit_softcap: float = 0.0
    sliding_window: int = 0                       # 0 = all layers global
    sliding_pattern: int = 2                      # every ``sliding_pattern``-th layer is global (Gemma-2: 2, Gemma-3: 6), the others slide
    name: str = "transformer"

    @staticmethod
    def mixtral_8x7b() -> "TransformerConfig":
        return TransformerConfig(num_experts=8, num_experts_per_tok=2, name="mixtral-8x7b")

    @staticmethod
    def qwen3_8b() -> "TransformerConfig":
        return TransformerConfig(intermediate_size=12288, num_layers=36, vocab_size=151936, rms_eps=1e-6, qk_norm=True, name="qwen3-8b")

    @staticmethod
    def qwen3_30b_a3b() -> "TransformerConfig":
        return TransformerConfig(hidden_size=2048, intermediate_size=768, num_layers=48, num_kv_heads=4, vocab_size=151936, rms_eps=1e-6,
                                 qk_norm=True, num_experts=128, num_experts_per_tok=8, name="qwen3-30b-a3b")

    @staticmethod
    def gemma2_9b() -> "TransformerConfig":
        return TransformerConfig(hidden_size=3584, intermediate_size=14336, num_layers=42, num_qo_heads=16, num_kv_heads=8, head_dim=256,
                                 vocab_size=256000, rms_eps=1e-6, rope_theta=1e4, activation="gelu_tanh", gemma_norm=True, post_norms=True,
                                 embed_scale=math.sqrt(3584.0), query_pre_attn_scalar=256.0, attn_logit_softcap=50.0, final_logit_softcap=30.0,
                                 sliding_window=4096, sliding_pattern=2, name="gemma-2-9b")

    def tiny(self) -> "TransformerConfig":
        """The same family switches at test size."""
        t = TransformerConfig(**self.__dict__)
        t.hidden_size, t.intermediate_size, t.num_layers, t.num_qo_heads, t.num_kv_heads, t.head_dim, t.vocab_size = 128, 192, 3, 4, 2, 64, 320
        if t.num_experts:
            t.num_experts, t.num_experts_per_tok, t.intermediate_size = 4, 2, 64
        if t.sliding_window:
            t.sliding_window = 6
        if t.embed_scale != 1.0:
            t.embed_scale = math.sqrt(t.hidden_size)
        if t.query_pre_attn_scalar:
            t.query_pre_attn_scalar = float(t.head_dim)
        t.name += "-tiny"
        return t

    @property
    def softmax_scale(self) -> float:
        return (self.query_pre_attn_scalar or self.head_dim) ** -0.5

    def is_sliding(self, layer: int) -> bool:
        return bool(self.sliding_window) and (layer + 1) % self.sliding_pattern != 0


class TransformerDecodeEngine:
    """Random-init decoder: batched single-token decode over a paged KV cache (``plan`` / ``step`` like the other engines)."""

    def __init__(self, cfg: TransformerConfig, max_batch: int, max_pages: int, page_size: int = 16, device: str = "cuda",
                 dtype: torch.dtype = torch.bfloat16, seed: int = 0, tp_group=None) -> None:
        """``tp_group``: tensor / expert parallelism over a ``torch.distributed`` group - attention heads, MLP columns and whole experts
        are sharded over its ranks (weights are drawn for the full model from ``seed`` and sliced, so every world size computes the same
        function); the two partial sums per layer are all-reduced (NCCL over NVLink on CUDA, gloo on CPU)."""
        self.cfg, self.page_size, self.max_batch = cfg, page_size, max_batch
        self.device, self.dtype = torch.device(device), dtype
        self.tp_group = tp_group
        self.tp_size = torch.distributed.get_world_size(tp_group) if tp_group is not None else 1
        self.tp_rank = torch.distributed.get_rank(tp_group) if tp_group is not None else 0
        tp, rk = self.tp_size, self.tp_rank
        if cfg.num_qo_heads % tp or cfg.num_kv_heads % tp or (cfg.num_experts or tp) % tp or cfg.intermediate_size % tp:
            raise ValueError(f"heads ({cfg.num_qo_heads} / {cfg.num_kv_heads}), experts and the MLP width must be divisible by the TP size {tp}")
        g = torch.Generator(device="cpu").manual_seed(seed)

        def w(rows: int, cols: int) -> torch.Tensor:
            return (torch.randn(rows, cols, generator=g) / cols ** 0.5).to(dtype).to(self.device)

        def norm_w(n: int) -> torch.Tensor:
            base = 0.1 * torch.randn(n, generator=g)
            return (base if cfg.gemma_norm else 1.0 + base).to(dtype).to(self.device)

        h, d, hq, hkv = cfg.hidden_size, cfg.head_dim, cfg.num_qo_heads, cfg.num_kv_heads
        self.hq, self.hkv = hq // tp, hkv // tp                      # local head counts
        self.embed = (torch.randn(cfg.vocab_size, h, generator=g) * 0.5).to(dtype).to(self.device)
        self.lm_head = w(cfg.vocab_size, h)
        self.final_norm = norm_w(h)
        self.layers: List[dict] = []

        rows = lambda t, blocks: shard_rows(t, rk, tp, blocks)  # noqa: E731

        for _ in range(cfg.num_layers):
            wq, wk, wv = w(hq * d, h), w(hkv * d, h), w(hkv * d, h)
            wo = w(h, hq * d)
            l = {"ln1": norm_w(h), "ln2": norm_w(h), "wqkv": torch.cat([rows(wq, 1), rows(wk, 1), rows(wv, 1)]).contiguous(),
                 "wo": shard_cols(wo, rk, tp),
                 "k_cache": torch.zeros(max_pages, page_size, self.hkv, d, dtype=dtype, device=self.device),
                 "v_cache": torch.zeros(max_pages, page_size, self.hkv, d, dtype=dtype, device=self.device)}
            if cfg.qk_norm:
                l.update(q_norm=norm_w(d), k_norm=norm_w(d))
            if cfg.post_norms:
                l.update(post_attn=norm_w(h), post_ffn=norm_w(h))
            if cfg.num_experts:
                e, i = cfg.num_experts, cfg.intermediate_size
                el = e // tp                                           # expert parallel: whole experts per rank
                w1 = (torch.randn(e, 2 * i, h, generator=g) / h ** 0.5).to(dtype)                                    # rows = [up | gate]
                w2 = (torch.randn(e, h, i, generator=g) / i ** 0.5).to(dtype)
                l.update(router=(torch.randn(e, h, generator=g) / h ** 0.5).to(dtype).to(self.device),
                         w1=w1[rk * el:(rk + 1) * el].contiguous().to(self.device), w2=w2[rk * el:(rk + 1) * el].contiguous().to(self.device))
            else:
                w_gu, w_d = w(2 * cfg.intermediate_size, h), w(h, cfg.intermediate_size)                             # rows = [gate | up]
                l.update(w_gu=rows(w_gu, 2), w_d=shard_cols(w_d, rk, tp))
            self.layers.append(l)
        ws = lambda: torch.empty(32 << 20, dtype=torch.uint8, device=self.device)  # noqa: E731
        self.attn_global = BatchDecodeWithPagedKVCacheWrapper(ws(), "NHD")
        self.attn_sliding = BatchDecodeWithPagedKVCacheWrapper(ws(), "NHD") if cfg.sliding_window else None
        self.prefill_global = BatchPrefillWithPagedKVCacheWrapper(ws(), "NHD")
        self.prefill_sliding = BatchPrefillWithPagedKVCacheWrapper(ws(), "NHD") if cfg.sliding_window else None
        self._mode = "decode"
        self.logits: Optional[torch.Tensor] = None

    def _norm(self, x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        return (norm.gemma_rmsnorm if self.cfg.gemma_norm else norm.rmsnorm)(x, weight, self.cfg.rms_eps)

    def plan(self, kv_indptr: torch.Tensor, kv_indices: torch.Tensor, kv_last_page_len: torch.Tensor) -> None:
        """``kv_*`` describe the cache INCLUDING the token this step appends."""
        cfg = self.cfg
        self.batch = b = kv_last_page_len.numel()
        if b > self.max_batch:
            raise ValueError(f"batch {b} exceeds max_batch {self.max_batch}")
        self.kv_indptr = kv_indptr.to(self.device, torch.int32)
        self.kv_indices = kv_indices.to(self.device, torch.int32)
        self.kv_last = kv_last_page_len.to(self.device, torch.int32)
        self.positions = (page.get_seq_lens(self.kv_indptr, self.kv_last, self.page_size).int() - 1).contiguous()
        self.batch_indices = torch.arange(b, device=self.device, dtype=torch.int32)
        common = dict(q_data_type=self.dtype, sm_scale=cfg.softmax_scale, logits_soft_cap=cfg.attn_logit_softcap or None)
        self.attn_global.plan(kv_indptr, kv_indices, kv_last_page_len, self.hq, self.hkv, cfg.head_dim, self.page_size, **common)
        if self.attn_sliding is not None:                       # a window of W tokens = the query plus W - 1 tokens to its left
            self.attn_sliding.plan(kv_indptr, kv_indices, kv_last_page_len, self.hq, self.hkv, cfg.head_dim, self.page_size,
                                   window_left=cfg.sliding_window - 1, **common)
        self.tokens = torch.zeros(b, dtype=torch.int64, device=self.device)
        self.next_tokens = torch.zeros(b, dtype=torch.int64, device=self.device)

    def _attention(self, li: int, l: dict, x: torch.Tensor) -> torch.Tensor:
        """``x [rows, hidden]``: one row per request in decode, all prompt tokens in prefill (``self._mode``)."""
        cfg, n = self.cfg, x.shape[0]
        hq, hkv, d = self.hq, self.hkv, cfg.head_dim
        qkv = linear(x, l["wqkv"]).view(n, hq + 2 * hkv, d)
        q, k, v = qkv[:, :hq].contiguous(), qkv[:, hq:hq + hkv].contiguous(), qkv[:, hq + hkv:].contiguous()
        if cfg.qk_norm:
            q = self._norm(q.view(n * hq, d), l["q_norm"]).view(n, hq, d)
            k = self._norm(k.view(n * hkv, d), l["k_norm"]).view(n, hkv, d)
        q, k = rope.apply_rope_pos_ids(q, k, self.positions, rope_theta=cfg.rope_theta)
        page.append_paged_kv_cache(k, v, self.batch_indices, self.positions, (l["k_cache"], l["v_cache"]), self.kv_indices, self.kv_indptr,
                                   self.kv_last, "NHD")
        if self._mode == "prefill":
            wrapper = self.prefill_sliding if cfg.is_sliding(li) else self.prefill_global
        else:
            wrapper = self.attn_sliding if cfg.is_sliding(li) else self.attn_global
        o = wrapper.run(q, (l["k_cache"], l["v_cache"]))
        return self._all_reduce(linear(o.reshape(n, hq * d), l["wo"]))

    def _ffn(self, l: dict, x: torch.Tensor) -> torch.Tensor:
        cfg = self.cfg
        if cfg.num_experts:
            logits = linear(x, l["router"])
            ids, wts = route(logits, None, cfg.num_experts_per_tok, int(RoutingMethodType.Renormalize))
            local = cfg.num_experts // self.tp_size
            return self._all_reduce(moe_forward(x, ids, wts, l["w1"], l["w2"], self.tp_rank * local, cfg.num_experts))
        act = activation.silu_and_mul if cfg.activation == "silu" else activation.gelu_tanh_and_mul
        return self._all_reduce(linear(act(linear(x, l["w_gu"])), l["w_d"]))

    def _all_reduce(self, partial: torch.Tensor) -> torch.Tensor:
        return all_reduce_fp32(partial, self.tp_group)

    def _add_norm(self, delta: torch.Tensor, res: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        """``res += delta`` and the RMSNorm of the new residual, in one kernel (in place on both; returns the normalised tensor)."""
        (norm.gemma_fused_add_rmsnorm if self.cfg.gemma_norm else norm.fused_add_rmsnorm)(delta, res, weight, self.cfg.rms_eps)
        return delta

    def _forward(self, tokens: torch.Tensor) -> torch.Tensor:
        """Final-norm-ed hidden states for ``tokens`` (rows laid out as the current plan says).  Every residual update is fused with the
        RMSNorm that follows it (the next block's input norm, at the end the final norm)."""
        cfg = self.cfg
        res = self.embed[tokens]
        if cfg.embed_scale != 1.0:
            res = (res.float() * cfg.embed_scale).to(self.dtype)
        x = self._norm(res, self.layers[0]["ln1"])
        for li, l in enumerate(self.layers):
            a = self._attention(li, l, x)
            if cfg.post_norms:
                a = self._norm(a, l["post_attn"])
            x = self._add_norm(a, res, l["ln2"])
            f = self._ffn(l, x)
            if cfg.post_norms:
                f = self._norm(f, l["post_ffn"])
            x = self._add_norm(f, res, self.layers[li + 1]["ln1"] if li + 1 < len(self.layers) else self.final_norm)
        return x

    def _head(self, hidden: torch.Tensor) -> torch.Tensor:
        cfg = self.cfg
        logits = linear(hidden, self.lm_head)                      # ``hidden`` already carries the final norm
        if cfg.final_logit_softcap:
            logits = (torch.tanh(logits.float() / cfg.final_logit_softcap) * cfg.final_logit_softcap).to(logits.dtype)
        return logits

    def step(self) -> torch.Tensor:
        self._mode = "decode"
        self.logits = self._head(self._forward(self.tokens))
        torch.argmax(self.logits, dim=-1, out=self.next_tokens)
        return self.next_tokens

    # ------------------------------------------------------------------ prompt processing
    def prefill(self, tokens: torch.Tensor, qo_indptr: torch.Tensor, kv_indptr: torch.Tensor, kv_indices: torch.Tensor,
                kv_last_page_len: torch.Tensor, all_logits: bool = False) -> torch.Tensor:
        """Append the packed prompt ``tokens`` (``qo_indptr [B + 1]`` splits them per request) to the paged cache described by ``kv_*``
        (lengths INCLUDE these tokens; earlier cache contents are the prefix) and run causal attention over prefix + prompt.  Returns the
        greedy next token of every request; ``self.logits`` holds the logits of each request's last prompt token (``all_logits``: of every
        appended token - what speculative decoding verifies drafts against)."""
        cfg = self.cfg
        self._mode = "prefill"
        b = kv_last_page_len.numel()
        self.kv_indptr = kv_indptr.to(self.device, torch.int32)
        self.kv_indices = kv_indices.to(self.device, torch.int32)
        self.kv_last = kv_last_page_len.to(self.device, torch.int32)
        seq_lens = page.get_seq_lens(self.kv_indptr, self.kv_last, self.page_size).int()
        qo = qo_indptr.to(self.device, torch.int32)
        self.batch_indices, self.positions = page.get_batch_indices_positions(qo, seq_lens, int(tokens.numel()))
        common = dict(causal=True, q_data_type=self.dtype, sm_scale=cfg.softmax_scale, logits_soft_cap=cfg.attn_logit_softcap or None)
        self.prefill_global.plan(qo_indptr, kv_indptr, kv_indices, kv_last_page_len, self.hq, self.hkv, cfg.head_dim,
                                 self.page_size, **common)
        if self.prefill_sliding is not None:
            self.prefill_sliding.plan(qo_indptr, kv_indptr, kv_indices, kv_last_page_len, self.hq, self.hkv, cfg.head_dim,
                                      self.page_size, window_left=cfg.sliding_window - 1, **common)
        hidden = self._forward(tokens.to(self.device))
        last = (qo[1:] - 1).long()
        self.logits = self._head(hidden if all_logits else hidden[last])      # all_logits: one row per appended token (verification of drafts)
        self._mode = "decode"
        return torch.argmax(self.logits[last] if all_logits else self.logits, dim=-1)
