"""Attention backends behind :class:`~flashinfer_b200.parallel_attention.ParallelAttention` (reference
flashinfer/parallel_attention/attention_ops.py).

A backend is a callable ``impl(query, key, value, attn_mask=None, is_causal=False, return_lse=False, tensor_layout="HND",
cur_rank_cu_seqlens_q=None, cur_rank_cu_seqlens_k=None, cur_rank_max_seqlen_q=0, cur_rank_max_seqlen_k=0, **kwargs)`` returning the
output in the input layout, plus - with ``return_lse`` - the NATURAL-log softmax denominator shaped like the output without its last
dimension (``[S, H]`` for NHD, ``[H, S]`` for HND).  The ring wrapper merges partial results with that statistic.

One native backend, registered as ``"sm100"`` and under the reference's names ``"cutlass"`` and ``"flash-attn3"`` (code written for the
reference selects one of those): the tcgen05 prefill kernel through ``single_prefill_with_kv_cache`` for one sequence and through the
ragged batch wrapper for packed (varlen) shards."""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import torch

from .utils import convert_output_layout, convert_qkv_layout

_LN2 = math.log(2.0)


class AttentionOpManager:
    """Registry of attention backends (``register_attn`` decorator, ``get_impl`` returns a fresh instance)."""
    _attn_registry: Dict[str, type] = {}
    attn_type: str = "sm100"

    @classmethod
    def op_type(cls) -> str:
        return "attention"

    @classmethod
    def set_attn_config(cls, **kwargs) -> None:
        for key, value in kwargs.items():
            if not hasattr(cls, key):
                raise AttributeError(f"'{cls.__name__}' has no attribute '{key}'")
            setattr(cls, key, value)

    @classmethod
    def register_attn(cls, attn_type: str) -> Callable[[type], type]:
        def decorator(attn_class: type) -> type:
            cls._attn_registry[attn_type] = attn_class
            return attn_class

        return decorator

    @classmethod
    def get_impl(cls, name: Optional[str] = None):
        name = name or cls.attn_type
        attn_class = cls._attn_registry.get(name)
        if attn_class is None:
            raise ValueError(f"Attention function {name} not found in registry (registered: {cls.get_registered_types()})")
        return attn_class()

    @classmethod
    def get_registered_types(cls) -> List[str]:
        return list(cls._attn_registry)


@AttentionOpManager.register_attn("sm100")
class NativeSm100Attention:
    """The library's own prefill kernels.  ``sm_scale`` (keyword) overrides the default ``1 / sqrt(head_dim)``."""

    def __call__(self, query, key, value, attn_mask=None, is_causal=False, return_lse=False, tensor_layout="HND", cur_rank_cu_seqlens_q=None,
                 cur_rank_cu_seqlens_k=None, cur_rank_max_seqlen_q=0, cur_rank_max_seqlen_k=0, sm_scale: Optional[float] = None, **kwargs):
        from ..prefill import BatchPrefillWithRaggedKVCacheWrapper, single_prefill_with_kv_cache

        if attn_mask is not None:
            raise NotImplementedError("attn_mask is not supported by the parallel attention backends")
        q, k, v = convert_qkv_layout(query, key, value, tensor_layout, "NHD")
        origin = q.dtype
        if q.is_cuda and q.dtype not in (torch.float16, torch.bfloat16):
            q, k, v = (t.to(torch.float16) for t in (q, k, v))
        if cur_rank_cu_seqlens_q is None:
            out, lse = single_prefill_with_kv_cache(q, k, v, causal=is_causal, sm_scale=sm_scale, return_lse=True)
        else:
            cq, ck = cur_rank_cu_seqlens_q.to(torch.int32), cur_rank_cu_seqlens_k.to(torch.int32)
            nq = int(cq[-1])
            out = q.new_zeros(q.shape[0], q.shape[1], v.shape[-1])
            lse = torch.full((q.shape[0], q.shape[1]), float("-inf"), dtype=torch.float32, device=q.device)   # rows outside every sequence
            if nq > 0:
                w = BatchPrefillWithRaggedKVCacheWrapper(torch.empty(32 << 20, dtype=torch.uint8, device=q.device), "NHD")
                w.plan(cq, ck, q.shape[1], k.shape[1], q.shape[2], head_dim_vo=v.shape[2], causal=is_causal, sm_scale=sm_scale, q_data_type=q.dtype)
                o, l = w.run(q[:nq].contiguous(), k[: int(ck[-1])].contiguous(), v[: int(ck[-1])].contiguous(), return_lse=True)
                out[:nq], lse[:nq] = o, l
        lse = lse.float() * _LN2                          # the kernels report base-2 statistics
        out = convert_output_layout(out.to(origin), "NHD", tensor_layout)
        if tensor_layout == "HND":
            lse = lse.transpose(0, 1).contiguous()
        return (out, lse) if return_lse else out


# the reference's backend names select the same native kernels here
AttentionOpManager.register_attn("cutlass")(NativeSm100Attention)
AttentionOpManager.register_attn("flash-attn3")(NativeSm100Attention)
CutlassFmha = FlashAttn3 = NativeSm100Attention
