"""Llama-family decode engine built only from this library's ops (flagship serving step).

One process per GPU; tensor parallelism shards attention heads and MLP columns over ``tp_size``
ranks.  With ``tp_size > 1`` the two per-layer reductions run through the in-kernel NVLink
all-reduce fused with residual-add + RMSNorm (``flashinfer_b200.comm``) — no NCCL on that path.
The whole step is launch-only (plan() is done once per batch composition), so it is captured in a
CUDA graph by ``LlamaDecodeEngine.capture()``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import torch

from .. import activation, norm, page, rope
from ..decode import BatchDecodeWithPagedKVCacheWrapper
from ..gemm.decode_linear import (EPI_GATED_SILU, EPI_RESIDUAL, EPI_ROPE_APPEND, FusedLinearTP, decode_linear, decode_prep,
                                  fold_rmsnorm_weight, permute_rope_rows, to_block_major_k)
from ..gemm.dense import interleave_gate_up, linear, linear_gated_silu


@dataclass
class LlamaConfig:
    hidden_size: int = 4096
    intermediate_size: int = 14336
    num_layers: int = 32
    num_qo_heads: int = 32
    num_kv_heads: int = 8
    head_dim: int = 128
    vocab_size: int = 128256
    rms_eps: float = 1e-5
    rope_theta: float = 5e5
    rope_scale: float = 8.0
    name: str = "llama-3-8b"

    @staticmethod
    def llama3_8b() -> "LlamaConfig":
        return LlamaConfig()

    @staticmethod
    def llama3_70b() -> "LlamaConfig":
        return LlamaConfig(hidden_size=8192, intermediate_size=28672, num_layers=80, num_qo_heads=64, num_kv_heads=8,
                           name="llama-3-70b")

    @staticmethod
    def tiny() -> "LlamaConfig":
        return LlamaConfig(hidden_size=512, intermediate_size=1024, num_layers=2, num_qo_heads=8, num_kv_heads=2,
                           vocab_size=1024, name="llama-tiny")


class LlamaDecodeEngine:
    """Random-init Llama decoder running batched single-token decode over a paged KV cache."""

    def __init__(self, cfg: LlamaConfig, max_batch: int, max_pages: int, page_size: int = 16, tp_rank: int = 0,
                 tp_size: int = 1, device: str = "cuda", dtype: torch.dtype = torch.bfloat16, comm=None, seed: int = 0,
                 fused: Optional[bool] = None, random_norms: bool = False, tp_group=None, kv_layout: str = "NHD",
                 block_major_k: Optional[bool] = None):
        """``fused`` (default: on for batches of at most 64 tokens): five launches per layer through
        :mod:`flashinfer_b200.gemm.decode_linear` - RMSNorm folded into the QKV / gate-up weights, RoPE + paged-KV append in the
        QKV epilogue, SwiGLU in the gate-up epilogue, residual add + norm statistics (+ the tensor-parallel all-reduce over
        NVLink) in the O / down epilogues.  ``fused=False`` keeps the op-by-op composition (any batch size)."""
        self.cfg, self.tp_rank, self.tp_size = cfg, tp_rank, tp_size
        self.device, self.dtype = torch.device(device), dtype
        self.page_size, self.max_batch = page_size, max_batch
        self.comm = comm
        self.fused = (max_batch <= 64 and cfg.head_dim % 32 == 0) if fused is None else bool(fused)
        if kv_layout not in ("NHD", "HND"):
            raise ValueError("kv_layout must be NHD or HND")
        self.kv_layout = kv_layout
        # BlockMajorK weights ([K / 64, N, 64]: every TMA box of a weight tile is one contiguous chunk of HBM) for the fused path
        self.block_major_k = (self.fused and self.device.type == "cuda") if block_major_k is None else bool(block_major_k)
        if self.fused and max_batch > 64:
            raise ValueError("the fused decode path handles at most 64 tokens per step")
        self.kv_prefetch = os.environ.get("FIB200_KV_PREFETCH", "1") != "0"
        self.tp_fused = None
        if self.fused and tp_size > 1:
            self.tp_fused = FusedLinearTP(tp_group if tp_group is not None else (comm.group if comm is not None else None),
                                          max_batch, cfg.hidden_size, dtype) if self.device.type == "cuda" else _CpuTP(tp_group)
        assert cfg.num_kv_heads % tp_size == 0 and cfg.intermediate_size % tp_size == 0
        self.hq = cfg.num_qo_heads // tp_size
        self.hkv = cfg.num_kv_heads // tp_size
        self.inter = cfg.intermediate_size // tp_size
        g = torch.Generator(device=self.device).manual_seed(seed + 1000 * tp_rank)
        gs = torch.Generator(device=self.device).manual_seed(seed)  # replicated tensors

        def rnd(shape, gen, std):
            return (torch.randn(shape, device=self.device, dtype=torch.float32, generator=gen) * std).to(dtype)

        h, d = cfg.hidden_size, cfg.head_dim
        self.embed = rnd((cfg.vocab_size, h), gs, 1.0)
        self.vocab_shard = (cfg.vocab_size + tp_size - 1) // tp_size
        self.lm_head = rnd((self.vocab_shard, h), g, h ** -0.5)
        def norm_w():
            if random_norms:  # tests: a non-trivial gain makes the folded and the op-by-op paths distinguishable
                return (1.0 + 0.2 * torch.randn(h, device=self.device, dtype=torch.float32, generator=gs)).to(dtype)
            return torch.ones(h, device=self.device, dtype=dtype)

        self.final_norm = norm_w()
        if self.fused:
            self.lm_head = fold_rmsnorm_weight(self.lm_head, self.final_norm)
            if self.block_major_k:
                self.lm_head = to_block_major_k(self.lm_head)
        self.layers = []
        cache_shape = (max_pages, page_size, self.hkv, d) if kv_layout == "NHD" else (max_pages, self.hkv, page_size, d)
        for _ in range(cfg.num_layers):
            l = {
                "ln1": norm_w(),
                "ln2": norm_w(),
                "wqkv": rnd(((self.hq + 2 * self.hkv) * d, h), g, h ** -0.5),
                "wo": rnd((h, self.hq * d), g, (cfg.num_qo_heads * d) ** -0.5),
                "wgu": interleave_gate_up(rnd((2 * self.inter, h), g, h ** -0.5)),  # (g0, u0, g1, u1, ...) rows
                "wd": rnd((h, self.inter), g, cfg.intermediate_size ** -0.5),
                "k_cache": torch.zeros(*cache_shape, device=self.device, dtype=dtype),
                "v_cache": torch.zeros(*cache_shape, device=self.device, dtype=dtype),
            }
            if self.fused:  # load-time weight preparation of the fused path (replaces the originals: no second copy)
                l["wqkv"] = permute_rope_rows(fold_rmsnorm_weight(l["wqkv"], l["ln1"]), self.hq, self.hkv, d)
                l["wgu"] = fold_rmsnorm_weight(l["wgu"], l["ln2"])
                if self.block_major_k:
                    for key in ("wqkv", "wo", "wgu", "wd"):
                        l[key] = to_block_major_k(l[key])
            self.layers.append(l)
        self._ws = torch.empty(64 << 20, dtype=torch.uint8, device=self.device)
        self.wrapper = BatchDecodeWithPagedKVCacheWrapper(self._ws, kv_layout)
        self.launches_per_step = 0
        self._graph = None

    # ------------------------------------------------------------------ planning
    def fill_kv_random(self, std: float = 0.5) -> None:
        for l in self.layers:
            l["k_cache"].normal_(0, std)
            l["v_cache"].normal_(0, std)

    def plan(self, kv_indptr: torch.Tensor, kv_indices: torch.Tensor, kv_last_page_len: torch.Tensor) -> None:
        """``kv_*`` describe the cache *including* the token that this step appends."""
        cfg = self.cfg
        self.batch = kv_last_page_len.numel()
        self.kv_indptr = kv_indptr.to(self.device, torch.int32)
        self.kv_indices = kv_indices.to(self.device, torch.int32)
        self.kv_last = kv_last_page_len.to(self.device, torch.int32)
        self.wrapper.plan(kv_indptr, kv_indices, kv_last_page_len, self.hq, self.hkv, cfg.head_dim, self.page_size,
                          q_data_type=self.dtype)
        seq_lens = page.get_seq_lens(self.kv_indptr, self.kv_last, self.page_size).int()
        self.positions = (seq_lens - 1).contiguous()
        self.batch_indices = torch.arange(self.batch, device=self.device, dtype=torch.int32)
        b, h = self.batch, cfg.hidden_size
        # static activations (CUDA-graph friendly)
        self.tokens = torch.zeros(b, dtype=torch.int64, device=self.device)
        self.next_tokens = torch.zeros(b, dtype=torch.int64, device=self.device)
        self._x = torch.empty(b, h, device=self.device, dtype=self.dtype)
        self._res = torch.empty(b, h, device=self.device, dtype=self.dtype)
        self._qkv = torch.empty(b, (self.hq + 2 * self.hkv) * cfg.head_dim, device=self.device, dtype=self.dtype)
        self._attn = torch.empty(b, self.hq, cfg.head_dim, device=self.device, dtype=self.dtype)
        self._attn_q = torch.empty(b, self.hq, cfg.head_dim, device=self.device, dtype=self.dtype)
        self._act = torch.empty(b, self.inter, device=self.device, dtype=self.dtype)
        self._logits = torch.empty(b, self.vocab_shard, device=self.device, dtype=self.dtype)
        if self.fused:
            self._sumsq = torch.zeros(2 * cfg.num_layers + 1, 64, device=self.device, dtype=torch.float32)
            self._cos_sin = torch.zeros(64, cfg.head_dim, device=self.device, dtype=torch.float32)
            self._cache_row = torch.zeros(64, device=self.device, dtype=torch.int64)
        self._graph = None

    # ------------------------------------------------------------------ one decode step
    def _row_parallel(self, inp: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
        """Row-parallel linear: partial sums go straight into the symmetric heap when tp > 1."""
        out = self._x if self.tp_size == 1 else self.comm.gemm_out(self.batch)
        return linear(inp, w, out=out)

    def _reduce_add_norm(self, part: torch.Tensor, weight: torch.Tensor) -> None:
        """self._x <- rmsnorm(residual += allreduce(part)); single-GPU: plain fused add+norm."""
        if self.tp_size == 1:
            norm.fused_add_rmsnorm(part, self._res, weight, self.cfg.rms_eps)
        else:
            self.comm.allreduce_add_rmsnorm(part, self._res, weight, self.cfg.rms_eps, out=self._x)

    def _step_fused(self) -> torch.Tensor:
        """Five launches per layer (see the class docstring); the residual stream ``self._res`` is the A operand of the
        QKV / gate-up GEMMs and is updated in place by the O / down GEMM epilogues."""
        cfg = self.cfg
        d, hq, hkv, h, b = cfg.head_dim, self.hq, self.hkv, cfg.hidden_size, self.batch
        res, ss = self._res, self._sumsq
        kc0 = self.layers[0]["k_cache"]
        slot_stride, head_stride = (kc0.stride(1), kc0.stride(2)) if self.kv_layout == "NHD" else (kc0.stride(2), kc0.stride(1))
        decode_prep(self.tokens, self.embed, res, ss, self.positions, self.kv_indptr, self.kv_indices, self.page_size,
                    kc0.stride(0), slot_stride, self._cos_sin, self._cache_row, d, batch_indices=self.batch_indices,
                    rope_scale=cfg.rope_scale, rope_theta=cfg.rope_theta, llama31=(1.0, 4.0, 8192.0))
        q2d = self._attn_q.view(b, hq * d)
        for li, l in enumerate(self.layers):
            decode_linear(res, l["wqkv"], EPI_ROPE_APPEND, out=q2d, row_sumsq=ss[2 * li], norm_dim=h, eps=cfg.rms_eps,
                          cos_sin=self._cos_sin, cache_row=self._cache_row, k_cache=l["k_cache"], v_cache=l["v_cache"],
                          num_q_heads=hq, num_kv_heads=hkv, head_dim=d, head_stride=head_stride)
            # the QKV kernel just before only appended this step's token: older KV tiles are prefetched under its tail
            self.wrapper.run(self._attn_q, (l["k_cache"], l["v_cache"]), out=self._attn, kv_prefetch=self.kv_prefetch)
            decode_linear(self._attn.view(b, hq * d), l["wo"], EPI_RESIDUAL, residual=res, sumsq_out=ss[2 * li + 1],
                          tp=self.tp_fused)
            decode_linear(res, l["wgu"], EPI_GATED_SILU, out=self._act, row_sumsq=ss[2 * li + 1], norm_dim=h, eps=cfg.rms_eps)
            decode_linear(self._act, l["wd"], EPI_RESIDUAL, residual=res, sumsq_out=ss[2 * li + 2], tp=self.tp_fused)
        # LM head with the final RMSNorm folded in (weights prepared at load time)
        decode_linear(res, self.lm_head, out=self._logits, row_sumsq=ss[2 * len(self.layers)], norm_dim=h, eps=cfg.rms_eps)
        if self.tp_size == 1:
            if self._logits.is_cuda:
                from ..comm.allreduce import local_argmax  # one CTA of 1024 threads per row (torch.argmax: 45 us for [64, 128 K])

                local_argmax(self._logits, out=self.next_tokens)
            else:
                torch.argmax(self._logits, dim=-1, out=self.next_tokens)
        elif self.comm is not None and hasattr(self.comm, "argmax_logits"):
            # one kernel: shard argmax + (value, index) exchange over NVLink + winner selection
            self.comm.argmax_logits(self._logits, self.tp_rank * self.vocab_shard, out=self.next_tokens)
        else:
            val, idx = torch.max(self._logits.float(), dim=-1)
            self.next_tokens.copy_(self._argmax_gather(val, idx + self.tp_rank * self.vocab_shard))
        self.launches_per_step = 1 + 5 * len(self.layers) + 2
        return self.next_tokens

    def _argmax_gather(self, val: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        if self.comm is not None:
            return self.comm.argmax_gather(val, idx)
        import torch.distributed as dist  # CPU / gloo tests

        grp = self.tp_fused.group if self.tp_fused is not None else None
        vals = [torch.empty_like(val) for _ in range(self.tp_size)]
        idxs = [torch.empty_like(idx) for _ in range(self.tp_size)]
        dist.all_gather(vals, val, group=grp)
        dist.all_gather(idxs, idx, group=grp)
        best = torch.stack(vals, 0).argmax(0)
        return torch.stack(idxs, 0).gather(0, best[None])[0]

    def step(self) -> torch.Tensor:
        """tokens (self.tokens) -> next tokens (self.next_tokens); greedy sampling."""
        if self.fused:
            return self._step_fused()
        cfg = self.cfg
        d, hq, hkv = cfg.head_dim, self.hq, self.hkv
        x = self._x
        torch.index_select(self.embed, 0, self.tokens, out=x)
        self._res.zero_()
        n = 2
        first = True
        for li, l in enumerate(self.layers):
            if first:
                norm.fused_add_rmsnorm(x, self._res, l["ln1"], cfg.rms_eps)
                first = False
                n += 1
            linear(x, l["wqkv"], out=self._qkv)
            qkv = self._qkv.view(self.batch, hq + 2 * hkv, d)
            q, k, v = qkv[:, :hq], qkv[:, hq : hq + hkv], qkv[:, hq + hkv :]
            # one kernel: RoPE(q) in place, RoPE(k) and v straight into the cache pages
            rope.apply_rope_append_paged_kv_cache(q, k, v, self.positions, self.batch_indices, (l["k_cache"], l["v_cache"]),
                                                  self.kv_indices, self.kv_indptr, self.kv_layout, rope_scale=cfg.rope_scale,
                                                  rope_theta=cfg.rope_theta, llama31=(1.0, 4.0, 8192.0))
            self.wrapper.run(q, (l["k_cache"], l["v_cache"]), out=self._attn)
            part = self._row_parallel(self._attn.view(self.batch, hq * d), l["wo"])
            self._reduce_add_norm(part, l["ln2"])
            # gate / up GEMM with the SwiGLU in its epilogue (weights row-interleaved at load time): one kernel, no [B, 2I] tensor
            linear_gated_silu(x, l["wgu"], out=self._act)
            part = self._row_parallel(self._act, l["wd"])
            n += 7
            nxt = self.layers[li + 1]["ln1"] if li + 1 < len(self.layers) else self.final_norm
            self._reduce_add_norm(part, nxt)
            n += 1
        linear(x, self.lm_head, out=self._logits)
        n += 1
        if self.tp_size == 1:
            torch.argmax(self._logits, dim=-1, out=self.next_tokens)
        elif hasattr(self.comm, "argmax_logits"):
            self.comm.argmax_logits(self._logits, self.tp_rank * self.vocab_shard, out=self.next_tokens)
        else:
            val, idx = torch.max(self._logits.float(), dim=-1)
            self.next_tokens.copy_(self.comm.argmax_gather(val, idx + self.tp_rank * self.vocab_shard))
        n += 1
        self.launches_per_step = n
        return self.next_tokens

    # ------------------------------------------------------------------ CUDA graph
    def capture(self, warmup: int = 2) -> None:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self.step()

    def replay(self) -> torch.Tensor:
        if self._graph is None:
            return self.step()
        self._graph.replay()
        return self.next_tokens


class _CpuTP:
    """Process-group handle of the fused path's fp32 oracle (CPU tensors, gloo): the all-reduce of the residual epilogue."""

    def __init__(self, group):
        import torch.distributed as dist

        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
