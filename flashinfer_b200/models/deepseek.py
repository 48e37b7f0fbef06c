"""DeepSeek-V2 / V3 family decode engine: Multi-head Latent Attention in its absorbed form + routed mixture-of-experts with a shared
expert, built from this library's public ops (the model family behind BASELINE configuration #4: MLA decode + grouped-top-k MoE).

Per layer (one token per request, paged latent cache):

    x   = rmsnorm(h)
    q   = (q_b . rmsnorm(q_a . x))  or  (q_proj . x)                 -> per head [nope 128 | rope 64]
    ckv, k_pe = kv_a . x                                             -> latent 512 (rmsnorm-ed) and the shared rope key 64
    rope(q_pe, k_pe)  (interleaved pairs, like the checkpoints);  append (ckv, k_pe) to the paged latent cache
    q_abs = q_nope . W_UK                                            -> the key up-projection absorbed into the query: [H, 512]
    o_lat = MLA(q_abs, q_pe; ckv cache, kpe cache)                   -> tcgen05 MLA kernel: softmax(q_abs.ckv + q_pe.kpe) . ckv
    attn  = o_lat . W_UV                                             -> the value up-projection applied after attention: [H, 128]
    h   += o_proj . attn
    x    = rmsnorm(h)
    h   += dense SwiGLU MLP (first ``first_k_dense`` layers)   or   shared expert + sum_k w_k * expert_k(x)  with sigmoid scores, bias-corrected
            grouped top-k routing (``fused_topk_deepseek``) and the fused MoE pipeline (``moe_forward``)

Nothing here is device specific: on CUDA every op is a native kernel, on CPU the ops' eager paths run the same graph (that is how the
engine is tested against a plain non-absorbed PyTorch model).  The flagship bench keeps the Llama engine with its fused decode GEMM
family; this engine is op-by-op (13 launches + the MoE pipeline per layer) - fusing its small GEMMs the same way is future work."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch

from .. import activation, norm, page, rope
from ..fused_moe.core import fused_topk_deepseek, moe_forward
from ..gemm.dense import linear
from ..mla import BatchMLAPagedAttentionWrapper


@dataclass
class DeepSeekConfig:
    hidden_size: int = 7168
    num_layers: int = 61
    num_heads: int = 128
    q_lora_rank: Optional[int] = 1536
    kv_lora_rank: int = 512
    qk_nope_head_dim: int = 128
    qk_rope_head_dim: int = 64
    v_head_dim: int = 128
    intermediate_size: int = 18432                 # dense layers
    moe_intermediate_size: int = 2048
    num_experts: int = 256
    num_experts_per_tok: int = 8
    n_group: int = 8
    topk_group: int = 4
    routed_scaling_factor: float = 2.5
    num_shared_experts: int = 1
    first_k_dense: int = 3
    vocab_size: int = 129280
    rms_eps: float = 1e-6
    rope_theta: float = 10000.0
    name: str = "deepseek-v3"

    @staticmethod
    def deepseek_v3() -> "DeepSeekConfig":
        return DeepSeekConfig()

    @staticmethod
    def tiny() -> "DeepSeekConfig":
        """Same structure at test size (the MLA kernel's latent / rope widths are kept: 512 + 64)."""
        return DeepSeekConfig(hidden_size=256, num_layers=3, num_heads=4, q_lora_rank=96, intermediate_size=384, moe_intermediate_size=128,
                              num_experts=8, num_experts_per_tok=2, n_group=4, topk_group=2, first_k_dense=1, vocab_size=512,
                              name="deepseek-tiny")

    @property
    def qk_head_dim(self) -> int:
        return self.qk_nope_head_dim + self.qk_rope_head_dim

    @property
    def softmax_scale(self) -> float:
        return self.qk_head_dim ** -0.5


class DeepSeekDecodeEngine:
    """Random-init DeepSeek-style decoder: batched single-token decode over a paged latent (MLA) cache.

    ``plan(kv_indptr, kv_indices, kv_last_page_len)`` describes the cache INCLUDING the token this step appends (like
    :class:`~flashinfer_b200.models.llama.LlamaDecodeEngine`); ``step()`` consumes ``self.tokens`` and returns greedy next tokens;
    ``logits`` holds the last step's logits."""

    def __init__(self, cfg: DeepSeekConfig, max_batch: int, max_pages: int, page_size: int = 64, device: str = "cuda",
                 dtype: torch.dtype = torch.bfloat16, seed: int = 0) -> None:
        self.cfg, self.page_size, self.max_batch = cfg, page_size, max_batch
        self.device, self.dtype = torch.device(device), dtype
        g = torch.Generator(device="cpu").manual_seed(seed)

        def w(rows: int, cols: int) -> torch.Tensor:
            return (torch.randn(rows, cols, generator=g) / cols ** 0.5).to(dtype).to(self.device)

        def ones(n: int) -> torch.Tensor:
            return (1.0 + 0.1 * torch.randn(n, generator=g)).to(dtype).to(self.device)

        h, hq = cfg.hidden_size, cfg.num_heads
        self.embed = (torch.randn(cfg.vocab_size, h, generator=g) * 0.5).to(dtype).to(self.device)
        self.lm_head = w(cfg.vocab_size, h)
        self.final_norm = ones(h)
        self.layers: List[dict] = []
        for li in range(cfg.num_layers):
            l = {"ln1": ones(h), "ln2": ones(h), "kv_a": w(cfg.kv_lora_rank + cfg.qk_rope_head_dim, h), "kv_norm": ones(cfg.kv_lora_rank),
                 # kv_b of the checkpoints, split per head into the key and the value up-projection
                 "w_uk": (torch.randn(hq, cfg.qk_nope_head_dim, cfg.kv_lora_rank, generator=g) / cfg.kv_lora_rank ** 0.5).to(dtype).to(self.device),
                 "w_uv": (torch.randn(hq, cfg.v_head_dim, cfg.kv_lora_rank, generator=g) / cfg.kv_lora_rank ** 0.5).to(dtype).to(self.device),
                 "wo": w(h, hq * cfg.v_head_dim),
                 "ckv_cache": torch.zeros(max_pages, page_size, cfg.kv_lora_rank, dtype=dtype, device=self.device),
                 "kpe_cache": torch.zeros(max_pages, page_size, cfg.qk_rope_head_dim, dtype=dtype, device=self.device)}
            if cfg.q_lora_rank:
                l.update(q_a=w(cfg.q_lora_rank, h), q_norm=ones(cfg.q_lora_rank), q_b=w(hq * cfg.qk_head_dim, cfg.q_lora_rank))
            else:
                l.update(q_proj=w(hq * cfg.qk_head_dim, h))
            if li < cfg.first_k_dense:
                l.update(w_gu=w(2 * cfg.intermediate_size, h), w_d=w(h, cfg.intermediate_size))          # rows = [gate | up]
            else:
                e, i = cfg.num_experts, cfg.moe_intermediate_size
                l.update(router=(torch.randn(e, h, generator=g) / h ** 0.5).float().to(self.device),
                         router_bias=(torch.randn(e, generator=g) * 0.1).float().to(self.device),
                         w1=(torch.randn(e, 2 * i, h, generator=g) / h ** 0.5).to(dtype).to(self.device),          # rows = [up | gate]
                         w2=(torch.randn(e, h, i, generator=g) / i ** 0.5).to(dtype).to(self.device),
                         shared_gu=w(2 * i * cfg.num_shared_experts, h), shared_d=w(h, i * cfg.num_shared_experts))
            self.layers.append(l)
        self.attn = BatchMLAPagedAttentionWrapper(torch.empty(64 << 20, dtype=torch.uint8, device=self.device))
        self.logits: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------ batch description
    def plan(self, kv_indptr: torch.Tensor, kv_indices: torch.Tensor, kv_last_page_len: torch.Tensor) -> None:
        cfg = self.cfg
        self.batch = b = kv_last_page_len.numel()
        if b > self.max_batch:
            raise ValueError(f"batch {b} exceeds max_batch {self.max_batch}")
        self.kv_indptr = kv_indptr.to(self.device, torch.int32)
        self.kv_indices = kv_indices.to(self.device, torch.int32)
        self.kv_last = kv_last_page_len.to(self.device, torch.int32)
        seq_lens = page.get_seq_lens(self.kv_indptr, self.kv_last, self.page_size).int()
        self.positions = (seq_lens - 1).contiguous()
        self.batch_indices = torch.arange(b, device=self.device, dtype=torch.int32)
        self.attn.plan(torch.arange(b + 1, dtype=torch.int32), kv_indptr.to("cpu", torch.int32), kv_indices.to("cpu", torch.int32),
                       seq_lens.to("cpu"), cfg.num_heads, cfg.kv_lora_rank, cfg.qk_rope_head_dim, self.page_size, False, cfg.softmax_scale,
                       self.dtype, self.dtype)
        self.tokens = torch.zeros(b, dtype=torch.int64, device=self.device)
        self.next_tokens = torch.zeros(b, dtype=torch.int64, device=self.device)

    def prefill(self, tokens: torch.Tensor, qo_indptr: torch.Tensor, kv_indptr: torch.Tensor, kv_indices: torch.Tensor,
                kv_last_page_len: torch.Tensor, all_logits: bool = False) -> torch.Tensor:
        """Append the packed prompt ``tokens`` (split per request by ``qo_indptr``) to the latent cache described by ``kv_*`` (lengths
        INCLUDE these tokens) and run causal absorbed MLA over prefix + prompt.  Returns the greedy next token per request; ``self.logits``:
        logits of every request's last token (``all_logits``: of every appended token)."""
        cfg = self.cfg
        self.kv_indptr = kv_indptr.to(self.device, torch.int32)
        self.kv_indices = kv_indices.to(self.device, torch.int32)
        self.kv_last = kv_last_page_len.to(self.device, torch.int32)
        seq_lens = page.get_seq_lens(self.kv_indptr, self.kv_last, self.page_size).int()
        qo = qo_indptr.to(self.device, torch.int32)
        self.batch_indices, self.positions = page.get_batch_indices_positions(qo, seq_lens, int(tokens.numel()))
        self.attn.plan(qo_indptr.to("cpu", torch.int32), kv_indptr.to("cpu", torch.int32), kv_indices.to("cpu", torch.int32), seq_lens.to("cpu"),
                       cfg.num_heads, cfg.kv_lora_rank, cfg.qk_rope_head_dim, self.page_size, True, cfg.softmax_scale, self.dtype, self.dtype)
        hidden = self._forward(tokens.to(self.device))
        last = (qo[1:] - 1).long()
        self.logits = linear(hidden if all_logits else hidden[last], self.lm_head)
        return torch.argmax(self.logits[last] if all_logits else self.logits, dim=-1)

    # ------------------------------------------------------------------ one decode step
    def _attention(self, l: dict, x: torch.Tensor) -> torch.Tensor:
        cfg, b, hq = self.cfg, x.shape[0], self.cfg.num_heads          # b = rows: requests in decode, prompt tokens in prefill
        if cfg.q_lora_rank:
            q = linear(norm.rmsnorm(linear(x, l["q_a"]), l["q_norm"], cfg.rms_eps), l["q_b"])
        else:
            q = linear(x, l["q_proj"])
        q = q.view(b, hq, cfg.qk_head_dim)
        q_nope, q_pe = q[..., : cfg.qk_nope_head_dim], q[..., cfg.qk_nope_head_dim:].contiguous()
        kv = linear(x, l["kv_a"])
        ckv = norm.rmsnorm(kv[:, : cfg.kv_lora_rank].contiguous(), l["kv_norm"], cfg.rms_eps)
        k_pe = kv[:, cfg.kv_lora_rank:].reshape(b, 1, cfg.qk_rope_head_dim).contiguous()
        q_pe, k_pe = rope.apply_rope_pos_ids(q_pe, k_pe, self.positions, interleave=True, rope_theta=cfg.rope_theta)
        page.append_paged_mla_kv_cache(ckv, k_pe.view(b, cfg.qk_rope_head_dim), self.batch_indices, self.positions, l["ckv_cache"], l["kpe_cache"],
                                       self.kv_indices, self.kv_indptr, self.kv_last)
        # absorb W_UK into the query: [b, H, nope] x [H, nope, rank] -> [b, H, rank]
        q_abs = torch.bmm(q_nope.transpose(0, 1), l["w_uk"]).transpose(0, 1).contiguous()
        o_lat = self.attn.run(q_abs, q_pe.contiguous(), l["ckv_cache"], l["kpe_cache"])                    # [b, H, rank]
        attn = torch.bmm(o_lat.transpose(0, 1), l["w_uv"].transpose(1, 2)).transpose(0, 1)                 # [b, H, v_head_dim]
        return linear(attn.reshape(b, hq * cfg.v_head_dim), l["wo"])

    def _ffn(self, l: dict, x: torch.Tensor) -> torch.Tensor:
        cfg = self.cfg
        if "w_gu" in l:
            return linear(activation.silu_and_mul(linear(x, l["w_gu"])), l["w_d"])
        scores = x.float() @ l["router"].t()
        weights, ids = fused_topk_deepseek(scores, l["router_bias"], cfg.n_group, cfg.topk_group, cfg.num_experts_per_tok, cfg.routed_scaling_factor)
        routed = moe_forward(x, ids, weights, l["w1"], l["w2"], 0, cfg.num_experts)
        shared = linear(activation.silu_and_mul(linear(x, l["shared_gu"])), l["shared_d"])
        return routed + shared

    def _forward(self, tokens: torch.Tensor) -> torch.Tensor:
        """Final-norm-ed hidden states (every residual update is fused with the RMSNorm that follows it)."""
        cfg = self.cfg
        res = self.embed[tokens]                                        # residual stream [rows, hidden]
        x = norm.rmsnorm(res, self.layers[0]["ln1"], cfg.rms_eps)
        for li, l in enumerate(self.layers):
            a = self._attention(l, x)
            norm.fused_add_rmsnorm(a, res, l["ln2"], cfg.rms_eps)       # res += a ; a <- rmsnorm(res)
            f = self._ffn(l, a)
            nxt = self.layers[li + 1]["ln1"] if li + 1 < len(self.layers) else self.final_norm
            norm.fused_add_rmsnorm(f, res, nxt, cfg.rms_eps)            # res += f ; f <- rmsnorm(res) = next layer's input
            x = f
        return x

    def step(self) -> torch.Tensor:
        self.logits = linear(self._forward(self.tokens), self.lm_head)
        torch.argmax(self.logits, dim=-1, out=self.next_tokens)
        return self.next_tokens
