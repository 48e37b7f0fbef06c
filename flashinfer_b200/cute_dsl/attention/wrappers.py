"""``BatchPrefillCuteDSLWrapper`` / ``BatchMLADecodeCuteDSLWrapper`` / ``cute_dsl_mla_decode`` (reference
flashinfer/cute_dsl/attention/wrappers/{batch_prefill,batch_mla}.py): the call shapes of the reference's DSL wrappers on top of the
tcgen05 ragged prefill kernel and the tcgen05 MLA decode kernel of this library."""
from __future__ import annotations

from typing import Optional

import torch

from ...api_logging import flashinfer_api
from .variant import AttentionVariant, SigmoidAttention, sigmoid_attention_reference


class BatchPrefillCuteDSLWrapper:
    """Ragged (variable-length) FMHA: ``plan(qo_indptr, kv_indptr, heads, head dims, causal, sm_scale, dtypes, window_left, variant)``
    then ``run(q, k, v, out=None)`` with ``q [total_q, Hq, D]``, ``k / v [total_kv, Hkv, D]``.  Like the reference wrapper the default
    ``sm_scale`` is 1.0 (callers pass ``1 / sqrt(head_dim)``)."""

    @flashinfer_api
    def __init__(self, float_workspace_buffer: torch.Tensor, use_cuda_graph: bool = False) -> None:
        self._float_workspace_buffer = float_workspace_buffer
        self.device = float_workspace_buffer.device
        self._use_cuda_graph = use_cuda_graph
        self._inner = None
        self._variant: Optional[AttentionVariant] = None

    @flashinfer_api
    def plan(self, qo_indptr, kv_indptr, num_qo_heads, num_kv_heads, head_dim_qk, head_dim_vo=None, causal=True, sm_scale=1.0,
             q_data_type=torch.float16, kv_data_type=torch.float16, window_left: int = -1, variant: Optional[AttentionVariant] = None) -> None:
        from ...prefill import BatchPrefillWithRaggedKVCacheWrapper

        if head_dim_vo is not None and head_dim_vo != head_dim_qk and (head_dim_qk, head_dim_vo) != (192, 128):
            raise ValueError("head_dim_vo must equal head_dim_qk (or 192 / 128)")
        self._variant = variant
        self._plan_args = (qo_indptr, kv_indptr, bool(causal), float(sm_scale))
        if isinstance(variant, SigmoidAttention):
            self._inner = None            # non-softmax normalisation: see variant.SigmoidAttention
            return
        jit_args = None
        if variant is not None and variant.is_compiled_hook:
            uri = f"cute_dsl_{variant.name}_{str(q_data_type).replace('torch.', '')}_{head_dim_qk}_{head_dim_vo or head_dim_qk}"
            jit_args = variant.jit_args(uri, q_data_type, kv_data_type, head_dim_qk, head_dim_vo or head_dim_qk)
        ws = self._float_workspace_buffer
        self._inner = BatchPrefillWithRaggedKVCacheWrapper(ws.view(torch.uint8) if ws.dtype != torch.uint8 else ws, "NHD", jit_args=jit_args)
        self._inner.plan(qo_indptr, kv_indptr, num_qo_heads, num_kv_heads, head_dim_qk, head_dim_vo=head_dim_vo, causal=causal,
                         pos_encoding_mode="ALIBI" if (variant is not None and variant.alibi_slopes is not None) else "NONE",
                         window_left=window_left, logits_soft_cap=(variant.logits_soft_cap or None) if variant is not None else None,
                         sm_scale=sm_scale, q_data_type=q_data_type, kv_data_type=kv_data_type)
        if variant is not None and variant.alibi_slopes is not None:
            slopes = variant.alibi_slopes.to(self.device, torch.float32).contiguous()
            if slopes.numel() != num_qo_heads:
                raise ValueError("alibi_slopes must hold one slope per query head")
            self._inner._alibi = slopes   # caller-provided slopes instead of the default geometric schedule

    @flashinfer_api
    def run(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        v_ = self._variant
        if self._inner is None:
            if not isinstance(v_, SigmoidAttention):
                raise RuntimeError("Plan the prefill attention computation first!")
            qo, kv, causal, sm = self._plan_args
            res = sigmoid_attention_reference(q, k, v, qo, kv, causal, sm, v_)
            if out is not None:
                out.copy_(res)
                return out
            return res
        extra = v_.run_args() if (v_ is not None and v_.is_compiled_hook) else ()
        if v_ is None or v_.sinks is None:
            return self._inner.run(q, k, v, *extra, out=out)
        # ragged entry point: the sink joins the denominator through the (out, lse) state (the paged entry point folds it in-kernel)
        from ...attention._core import apply_attention_sink

        o, lse = self._inner.run(q, k, v, *extra, return_lse=True)
        o, _ = apply_attention_sink(o, lse, v_.sinks.to(o.device))
        if out is not None:
            out.copy_(o)
            return out
        return o


class BatchMLADecodeCuteDSLWrapper:
    """Absorbed MLA decode: ``plan(kv_lora_rank, qk_rope_head_dim, num_heads, page_size, q_dtype, ...)`` then
    ``run(q [B, q_len, H, 576], kv_cache [pages, (1,) page, 576], block_tables, seq_lens, max_seq_len, softmax_scale, output_scale)``."""

    @flashinfer_api
    def __init__(self, workspace_buffer: torch.Tensor) -> None:
        if workspace_buffer.dtype not in (torch.int8, torch.uint8):
            raise TypeError(f"workspace_buffer must be a byte tensor, got {workspace_buffer.dtype}")
        self._workspace_buffer = workspace_buffer.view(torch.uint8)
        self._device = workspace_buffer.device
        self._planned = False

    @flashinfer_api
    def plan(self, kv_lora_rank: int = 512, qk_rope_head_dim: int = 64, num_heads: int = 128, page_size: int = 1,
             q_dtype: torch.dtype = torch.bfloat16, out_dtype: Optional[torch.dtype] = None, is_var_seq: bool = True,
             enable_pdl: Optional[bool] = None, variant: Optional[AttentionVariant] = None) -> None:
        if variant is not None and (variant.is_compiled_hook or variant.alibi_slopes is not None or variant.logits_soft_cap
                                    or variant.sinks is not None):
            raise NotImplementedError("the MLA decode kernel has no variant hooks (prefill kernel only)")
        if q_dtype not in (torch.float16, torch.bfloat16):
            raise NotImplementedError("MLA decode: fp16 / bf16 queries and cache (fp8 MLA is not implemented)")
        self._kv_lora_rank, self._qk_rope_head_dim, self._num_heads, self._page_size = kv_lora_rank, qk_rope_head_dim, num_heads, page_size
        self._q_dtype, self._o_dtype, self._enable_pdl = q_dtype, out_dtype or q_dtype, enable_pdl
        self._planned = True

    @flashinfer_api
    def run(self, q: torch.Tensor, kv_cache: torch.Tensor, block_tables: torch.Tensor, seq_lens: torch.Tensor, max_seq_len: int,
            softmax_scale: float, output_scale: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        from ...mla._core import trtllm_batch_decode_with_kv_cache_mla

        if not self._planned:
            raise RuntimeError("Plan the MLA decode computation first!")
        if q.shape[-1] != self._kv_lora_rank + self._qk_rope_head_dim or q.shape[-2] != self._num_heads:
            raise ValueError(f"q must be [B, q_len, {self._num_heads}, {self._kv_lora_rank + self._qk_rope_head_dim}], got {tuple(q.shape)}")
        res = trtllm_batch_decode_with_kv_cache_mla(q, kv_cache, self._workspace_buffer, 128, self._kv_lora_rank, self._qk_rope_head_dim,
                                                    block_tables, seq_lens, max_seq_len, out=out if (out is not None and out.dtype == q.dtype) else None,
                                                    bmm1_scale=softmax_scale, bmm2_scale=output_scale, enable_pdl=self._enable_pdl)
        if out is not None and res.data_ptr() != out.data_ptr():
            out.copy_(res)
            return out
        return res if res.dtype == self._o_dtype else res.to(self._o_dtype)


@flashinfer_api
def cute_dsl_mla_decode(query: torch.Tensor, kv_cache: torch.Tensor, workspace_buffer: torch.Tensor, kv_lora_rank: int, qk_rope_head_dim: int,
                        block_tables: torch.Tensor, seq_lens: torch.Tensor, max_seq_len: int, softmax_scale: float, output_scale: float = 1.0,
                        out: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None, is_var_seq: bool = True,
                        enable_pdl: Optional[bool] = None) -> torch.Tensor:
    """Function form of :class:`BatchMLADecodeCuteDSLWrapper` (plan + run in one call)."""
    w = BatchMLADecodeCuteDSLWrapper(workspace_buffer)
    page = kv_cache.shape[-2]
    w.plan(kv_lora_rank, qk_rope_head_dim, query.shape[-2], page, query.dtype, out_dtype, is_var_seq, enable_pdl)
    return w.run(query, kv_cache, block_tables, seq_lens, max_seq_len, softmax_scale, output_scale, out)
