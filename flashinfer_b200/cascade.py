"""Attention-state merge operators and cascade wrappers.  Parity: reference flashinfer/cascade.py:42-559."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import jit, reference
from .utils import dtype_code, stream_ptr


def merge_state(v_a, s_a, v_b, s_b) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge two attention states ``(v [n,H,D], s [n,H] base-2 lse)``."""
    if not v_a.is_cuda:
        return reference.merge_state_ref(v_a, s_a, v_b, s_b)
    v_a, v_b = v_a.contiguous(), v_b.contiguous()
    s_a, s_b = s_a.float().contiguous(), s_b.float().contiguous()
    v = torch.empty_like(v_a)
    s = torch.empty_like(s_a)
    n, h, d = v_a.shape
    jit.load("cascade").call("merge_state", v_a, s_a, v_b, s_b, v, s, None, n, h, d, dtype_code(v_a.dtype), 1,
                             stream_ptr(v_a))
    return v, s


def merge_state_in_place(v, s, v_other, s_other, mask: Optional[torch.Tensor] = None) -> None:
    """``(v, s) <- merge((v, s), (v_other, s_other))`` in place; rows with ``mask == False`` are untouched."""
    if not v.is_cuda:
        vo, so = reference.merge_state_ref(v, s, v_other, s_other)
        if mask is not None:
            m = mask.bool()
            v[m] = vo[m]
            s[m] = so[m]
        else:
            v.copy_(vo)
            s.copy_(so)
        return
    if not v.is_contiguous() or not s.is_contiguous() or s.dtype != torch.float32:
        raise ValueError("merge_state_in_place needs contiguous v and fp32 contiguous s")
    n, h, d = v.shape
    m8 = mask.to(torch.uint8).contiguous() if mask is not None else None
    jit.load("cascade").call("merge_state", v, s, v_other.contiguous(), s_other.float().contiguous(), v, s, m8, n, h, d,
                             dtype_code(v.dtype), 1, stream_ptr(v))


def merge_states(v: torch.Tensor, s: torch.Tensor, out: Optional[torch.Tensor] = None,
                 lse_out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """v [n, K, H, D], s [n, K, H] -> merged ([n, H, D], [n, H]).  Slots with ``s == -inf`` are ignored.  ``out`` (any
    16-bit / fp32 dtype matching ``v`` or not) and ``lse_out`` may be pre-allocated contiguous tensors."""
    if not v.is_cuda:
        vo, so = reference.merge_states_ref(v, s)
        if out is not None:
            out.copy_(vo)
            vo = out
        if lse_out is not None:
            lse_out.copy_(so)
            so = lse_out
        return vo, so
    v = v.contiguous()
    s = s.float().contiguous()
    n, k, h, d = v.shape
    direct = out is not None and out.dtype == v.dtype and out.is_contiguous()
    vo = out if direct else torch.empty(n, h, d, dtype=v.dtype, device=v.device)
    so = lse_out if (lse_out is not None and lse_out.is_contiguous() and lse_out.dtype == torch.float32) else \
        torch.empty(n, h, dtype=torch.float32, device=v.device)
    jit.load("cascade").call("merge_states", v, s, vo, so, n, k, h, d, dtype_code(v.dtype), 1, stream_ptr(v))
    if out is not None and not direct:
        out.copy_(vo)
        vo = out
    if lse_out is not None and so.data_ptr() != lse_out.data_ptr():
        lse_out.copy_(so)
        so = lse_out
    return vo, so


# ------------------------------------------------------------------------------------------------
# Cascade wrappers (reference flashinfer/cascade.py:226-1086): one prefill wrapper per level, every level
# returns (o, lse) and the levels are folded with the merge operator.
# ------------------------------------------------------------------------------------------------
class MultiLevelCascadeAttentionWrapper:
    """Multi-level cascade attention over a shared paged KV cache (level 0 = most shared prefix)."""

    def __init__(self, num_levels, float_workspace_buffer: torch.Tensor, kv_layout: str = "NHD",
                 use_cuda_graph: bool = False, qo_indptr_buf_arr=None, paged_kv_indptr_buf_arr=None,
                 paged_kv_indices_buf_arr=None, paged_kv_last_page_len_buf_arr=None) -> None:
        from .prefill import BatchPrefillWithPagedKVCacheWrapper

        self._num_levels = num_levels
        self._kv_layout = kv_layout
        self._use_cuda_graph = use_cuda_graph
        self._batch_prefill_wrappers = [
            BatchPrefillWithPagedKVCacheWrapper(float_workspace_buffer, kv_layout, use_cuda_graph)
            for _ in range(num_levels)
        ]

    @property
    def is_cuda_graph_enabled(self) -> bool:
        return self._use_cuda_graph

    def reset_workspace_buffer(self, float_workspace_buffer, int_workspace_buffers) -> None:
        for w, ib in zip(self._batch_prefill_wrappers, int_workspace_buffers):
            w.reset_workspace_buffer(float_workspace_buffer, ib)

    def plan(self, qo_indptr_arr, paged_kv_indptr_arr, paged_kv_indices_arr, paged_kv_last_page_len, num_qo_heads,
             num_kv_heads, head_dim, page_size, causal=False, pos_encoding_mode="NONE", use_fp16_qk_reduction=False,
             sm_scale=None, window_left=-1, logits_soft_cap=None, rope_scale=None, rope_theta=None,
             q_data_type="float16", kv_data_type=None) -> None:
        n = self._num_levels
        for i, (w, qo, kvp, kvi, last) in enumerate(zip(self._batch_prefill_wrappers, qo_indptr_arr, paged_kv_indptr_arr,
                                                        paged_kv_indices_arr, paged_kv_last_page_len)):
            # only the last (unique-suffix) level is causal, like the reference
            w.plan(qo, kvp, kvi, last, num_qo_heads, num_kv_heads, head_dim, page_size,
                   causal=causal if i == n - 1 else False, pos_encoding_mode=pos_encoding_mode, sm_scale=sm_scale,
                   window_left=window_left, logits_soft_cap=logits_soft_cap, q_data_type=q_data_type,
                   kv_data_type=kv_data_type)

    begin_forward = plan

    def run(self, q: torch.Tensor, paged_kv_cache):
        out, lse = self._batch_prefill_wrappers[-1].run(q, paged_kv_cache, return_lse=True)
        for w in self._batch_prefill_wrappers[:-1]:
            out_i, lse_i = w.run(q, paged_kv_cache, return_lse=True)
            merge_state_in_place(out, lse, out_i, lse_i)
        return out

    forward = run

    def end_forward(self) -> None:
        pass


class BatchDecodeWithSharedPrefixPagedKVCacheWrapper:
    """Decode with one prefix (contiguous k/v) shared by the whole batch + per-request paged suffixes."""

    def __init__(self, float_workspace_buffer: torch.Tensor, kv_layout: str = "NHD") -> None:
        from .decode import BatchDecodeWithPagedKVCacheWrapper

        self._batch_decode_wrapper = BatchDecodeWithPagedKVCacheWrapper(float_workspace_buffer, kv_layout)
        self._kv_layout = kv_layout

    def reset_workspace_buffer(self, float_workspace_buffer, int_workspace_buffer) -> None:
        self._batch_decode_wrapper.reset_workspace_buffer(float_workspace_buffer, int_workspace_buffer)

    def begin_forward(self, unique_kv_indptr, unique_kv_indices, unique_kv_last_page_len, num_qo_heads, num_kv_heads,
                      head_dim, page_size, data_type="float16") -> None:
        self._batch_decode_wrapper.plan(unique_kv_indptr, unique_kv_indices, unique_kv_last_page_len, num_qo_heads,
                                        num_kv_heads, head_dim, page_size, q_data_type=data_type)

    def forward(self, q, k_shared, v_shared, unique_kv_cache, allow_fp16_qk_reduction=False, sm_scale=None,
                rope_scale=None, rope_theta=None):
        from .prefill import single_prefill_with_kv_cache

        v_sh, s_sh = single_prefill_with_kv_cache(q, k_shared, v_shared, causal=False, kv_layout=self._kv_layout,
                                                  sm_scale=sm_scale, return_lse=True)
        v_un, s_un = self._batch_decode_wrapper.run(q, unique_kv_cache, return_lse=True)
        merge_state_in_place(v_sh, s_sh, v_un, s_un)
        return v_sh

    def end_forward(self) -> None:
        pass


class BatchPrefillWithSharedPrefixPagedKVCacheWrapper:
    """Prefill/append with one shared contiguous prefix + per-request paged suffixes."""

    def __init__(self, float_workspace_buffer: torch.Tensor, kv_layout: str = "NHD") -> None:
        from .prefill import BatchPrefillWithPagedKVCacheWrapper

        self._batch_prefill_wrapper = BatchPrefillWithPagedKVCacheWrapper(float_workspace_buffer, kv_layout)
        self._kv_layout = kv_layout

    def reset_workspace_buffer(self, float_workspace_buffer, int_workspace_buffer) -> None:
        self._batch_prefill_wrapper.reset_workspace_buffer(float_workspace_buffer, int_workspace_buffer)

    def begin_forward(self, qo_indptr, paged_kv_indptr, paged_kv_indices, paged_kv_last_page_len, num_qo_heads,
                      num_kv_heads, head_dim, page_size) -> None:
        self._plan_args = (qo_indptr, paged_kv_indptr, paged_kv_indices, paged_kv_last_page_len, num_qo_heads,
                           num_kv_heads, head_dim, page_size)

    def forward(self, q, k_shared, v_shared, unique_kv_cache, causal=False, allow_fp16_qk_reduction=False, sm_scale=None,
                rope_scale=None, rope_theta=None):
        from .prefill import single_prefill_with_kv_cache

        self._batch_prefill_wrapper.plan(*self._plan_args, causal=causal, sm_scale=sm_scale, q_data_type=q.dtype)
        v_sh, s_sh = single_prefill_with_kv_cache(q, k_shared, v_shared, causal=False, kv_layout=self._kv_layout,
                                                  sm_scale=sm_scale, return_lse=True)
        v_un, s_un = self._batch_prefill_wrapper.run(q, unique_kv_cache, return_lse=True)
        merge_state_in_place(v_sh, s_sh, v_un, s_un)
        return v_sh

    def end_forward(self) -> None:
        pass


def get_cascade_module(*args, **kwargs):
    """The native module behind this file's ops (reference cascade.py get_cascade_module: the JIT module accessor)."""
    from . import jit

    return jit.load("cascade")
