"""Prefill / append attention: single-request functions and the ragged / paged batch wrappers.

API parity: reference flashinfer/prefill.py:1111-1387 (single_prefill_with_kv_cache),
:1412-2536 (BatchPrefillWithPagedKVCacheWrapper), :2552-3555 (BatchPrefillWithRaggedKVCacheWrapper).

CUDA tensors run the persistent tcgen05 FMHA kernel (csrc/attention/prefill_sm100.cu) driven by the
C++ LPT planner; CPU tensors run the fp32 oracle (BASELINE.json config #1 is the CPU plumbing path).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple, Union

import torch

from . import jit, reference
from .utils import host_i32 as _host_i32
from .utils import legacy_forward_replan, remember_plan
from .utils import (
    check_kv_layout,
    check_pos_encoding_mode,
    device_sm_count,
    dtype_code,
    paged_kv_strides,
    stream_ptr,
    unpack_paged_kv_cache,
)

_WORK_INTS = 8
_TILE_Q = 256
_TILE_KV = 128


def _canon_dtype(dt):
    return getattr(torch, dt) if isinstance(dt, str) else dt


def _unpack_bits(packed: torch.Tensor, n: int) -> torch.Tensor:
    """Little-endian bit unpack (inverse of quantization.packbits)."""
    shifts = torch.arange(8, device=packed.device, dtype=torch.uint8)
    bits = ((packed.view(torch.uint8)[:, None] >> shifts[None, :]) & 1).flatten()[:n]
    return bits.bool()


def _unpack_segmented(packed: torch.Tensor, seg_bits) -> torch.Tensor:
    """Inverse of ``segment_packbits(mask, indptr, "little")`` (the format of ``packed_custom_mask`` in the reference: every
    request's ``q_len x kv_len`` bits start on a byte boundary, reference prefill.py ``_compute_page_mask_indptr``):
    concatenated bool bits of all segments."""
    out, off = [], 0
    flat = packed.reshape(-1)
    for n in seg_bits:
        nb = (int(n) + 7) // 8
        out.append(_unpack_bits(flat[off: off + nb], int(n)))
        off += nb
    return torch.cat(out) if out else torch.empty(0, dtype=torch.bool, device=packed.device)


def single_prefill_with_kv_cache(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    scale_q: Optional[torch.Tensor] = None,
    scale_k: Optional[torch.Tensor] = None,
    scale_v: Optional[torch.Tensor] = None,
    o_dtype: Optional[torch.dtype] = None,
    custom_mask: Optional[torch.Tensor] = None,
    packed_custom_mask: Optional[torch.Tensor] = None,
    causal: bool = False,
    kv_layout: str = "NHD",
    pos_encoding_mode: str = "NONE",
    use_fp16_qk_reduction: bool = False,
    sm_scale: Optional[float] = None,
    window_left: int = -1,
    logits_soft_cap: Optional[float] = None,
    rope_scale: Optional[float] = None,
    rope_theta: Optional[float] = None,
    backend: str = "auto",
    return_lse: bool = False,
    kv_cache_sf=None,
    k_scale: Optional[float] = None,
    v_scale: Optional[float] = None,
):
    """Prefill/append attention for one request.  q ``[qo_len, Hq, D]``; k/v ``[kv_len, Hkv, D]`` (NHD)
    or ``[Hkv, kv_len, D]`` (HND).  Returns ``o`` (and base-2 ``lse [qo_len, Hq]``).  ``k_scale`` folds into the softmax scale and
    ``v_scale`` multiplies the output (de-quantisation scales of a low-precision K / V); NVFP4 K / V (``kv_cache_sf``) is refused."""
    if kv_cache_sf is not None:
        raise NotImplementedError("single_prefill_with_kv_cache: NVFP4 K / V (kv_cache_sf) is not implemented")
    check_kv_layout(kv_layout)
    check_pos_encoding_mode(pos_encoding_mode)
    d = q.shape[-1]
    if sm_scale is None:
        sm_scale = 1.0 / math.sqrt(d)
    if k_scale is not None:
        sm_scale = sm_scale * float(k_scale)
    if kv_layout == "HND":
        k, v = k.transpose(0, 1), v.transpose(0, 1)
    qo_len, kv_len = q.shape[0], k.shape[0]
    if pos_encoding_mode == "ROPE_LLAMA":              # rotate q (the last qo_len positions) and k, then plain attention
        from .attention.rope_on_the_fly import rope_params, rotate_rows

        rs, rt = rope_params(rope_scale, rope_theta)
        q = rotate_rows(q, torch.arange(kv_len - qo_len, kv_len), rs, rt)
        k = rotate_rows(k.contiguous(), torch.arange(kv_len), rs, rt)
        pos_encoding_mode = "NONE"
    mask = None
    if packed_custom_mask is not None and custom_mask is None:
        mask = _unpack_bits(packed_custom_mask, qo_len * kv_len).view(qo_len, kv_len)
    elif custom_mask is not None:
        mask = custom_mask.view(qo_len, kv_len)
    fast = q.is_cuda and (q.shape[-1], v.shape[-1]) in ((128, 128), (192, 128), (64, 64)) and q.dtype in (
        torch.float16, torch.bfloat16) and k.dtype == q.dtype
    alibi = pos_encoding_mode == "ALIBI"
    if not q.is_cuda:
        from .utils import get_alibi_slopes

        o, lse = reference.attention_ref(q, k, v, causal and mask is None, sm_scale, logits_soft_cap or 0.0,
                                         window_left, custom_mask=mask, alibi_slopes=get_alibi_slopes(q.shape[1]) if alibi else None)
    elif not fast:
        from .attention import generic as _g

        if not _g.supported(q, k, q.shape[-1], v.shape[-1]):
            raise NotImplementedError(f"attention: unsupported configuration (dtype {q.dtype}/{k.dtype}, head_dim {q.shape[-1]})")
        o = torch.empty(qo_len, q.shape[1], v.shape[-1], dtype=q.dtype, device=q.device)
        lse = torch.empty(qo_len, q.shape[1], dtype=torch.float32, device=q.device)
        qi = torch.tensor([0, qo_len], dtype=torch.int32, device=q.device)
        ki = torch.tensor([0, kv_len], dtype=torch.int32, device=q.device)
        if k.stride(-1) != 1 or v.stride(-1) != 1:
            k, v = k.contiguous(), v.contiguous()
        pm = _g.pack_mask_bits(mask) if mask is not None else None
        mi = torch.tensor([0, qo_len * kv_len], dtype=torch.int32, device=q.device) if mask is not None else None
        ks = float(scale_k) if scale_k is not None else 1.0
        qs_ = float(scale_q) if scale_q is not None else 1.0
        from .utils import get_alibi_slopes

        _g.run(q, k, v, o, lse, qi, ki, None, None, 1, (0, k.stride(0), k.stride(1)), (0, v.stride(0), v.stride(1)), k.shape[1],
               causal and mask is None, window_left, sm_scale * qs_, logits_soft_cap or 0.0, pm, mi,
               get_alibi_slopes(q.shape[1], q.device) if alibi else None, ks, float(scale_v) if scale_v is not None else 1.0)
    else:
        ws = torch.empty(16 * 1024 * 1024, dtype=torch.uint8, device=q.device)
        w = BatchPrefillWithRaggedKVCacheWrapper(ws, "NHD")
        qo_indptr = torch.tensor([0, qo_len], dtype=torch.int32)
        kv_indptr = torch.tensor([0, kv_len], dtype=torch.int32)
        w.plan(qo_indptr, kv_indptr, q.shape[1], k.shape[1], d, head_dim_vo=v.shape[-1], causal=causal, sm_scale=sm_scale,
               window_left=window_left, logits_soft_cap=logits_soft_cap, q_data_type=q.dtype,
               custom_mask=mask.flatten() if mask is not None else None, pos_encoding_mode=pos_encoding_mode)
        o, lse = w.run(q, k, v, return_lse=True)
    if v_scale is not None and float(v_scale) != 1.0:
        o = (o.float() * float(v_scale)).to(o.dtype)
    if o_dtype is not None and o.dtype != o_dtype:
        o = o.to(o_dtype)
    return (o, lse) if return_lse else o


def single_prefill_with_kv_cache_return_lse(*args, **kwargs):
    kwargs["return_lse"] = True
    return single_prefill_with_kv_cache(*args, **kwargs)


class _BatchPrefillBase:
    """Shared plan()/run() machinery of the ragged and paged prefill wrappers."""

    _paged = False

    def __init__(self, float_workspace_buffer: torch.Tensor, kv_layout: str = "NHD", use_cuda_graph: bool = False,
                 qo_indptr_buf=None, paged_kv_indptr_buf=None, paged_kv_indices_buf=None,
                 paged_kv_last_page_len_buf=None, kv_indptr_buf=None, custom_mask_buf=None, mask_indptr_buf=None,
                 backend: str = "auto", jit_args=None, jit_kwargs=None):
        check_kv_layout(kv_layout)
        self._kv_layout = kv_layout
        self._float_workspace_buffer = float_workspace_buffer
        self.device = float_workspace_buffer.device
        self._use_cuda_graph = use_cuda_graph
        self._int_workspace_buffer = torch.empty(8 * 1024 * 1024, dtype=torch.uint8, device=self.device)
        self._pin_int_workspace_buffer = torch.empty(8 * 1024 * 1024, dtype=torch.uint8, device="cpu",
                                                     pin_memory=self.device.type == "cuda")
        self._planned = False
        self._backend = "sm100"
        self._cta_budget: Optional[int] = None  # POD: restrict the persistent grid to this many SMs
        # user attention variant (reference: jit_args of the wrappers + flashinfer/jit/attention/modules.py gen_customize_*):
        # LogitsTransform / LogitsMask hooks compiled INTO the tcgen05 prefill kernel (csrc/attention/prefill_sm100.cu)
        self._variant_mod = None
        self._variant_tensor_names, self._variant_scalar_names = [], []
        if jit_args is not None:
            spec = jit.gen_customize_batch_prefill_module(backend, *jit_args, **(jit_kwargs or {}))
            self._variant_mod = spec.build_and_load()
            self._variant_tensor_names = list(spec.additional_tensor_names)
            self._variant_scalar_names = list(spec.additional_scalar_names)
        self._alibi = None
        self._mask_words = self._mask_bit_indptr = None

    @property
    def is_cuda_graph_enabled(self) -> bool:
        return self._use_cuda_graph

    def reset_workspace_buffer(self, float_workspace_buffer, int_workspace_buffer) -> None:
        self._float_workspace_buffer = float_workspace_buffer
        self._int_workspace_buffer = int_workspace_buffer
        self._pin_int_workspace_buffer = torch.empty(int_workspace_buffer.numel(), dtype=torch.uint8, device="cpu",
                                                     pin_memory=self.device.type == "cuda")

    def _plan_common(self, qo_indptr, kv_lens_host, num_qo_heads, num_kv_heads, head_dim_qk, head_dim_vo, causal,
                     sm_scale, window_left, logits_soft_cap, q_data_type, kv_data_type, custom_mask, non_blocking):
        if num_qo_heads % num_kv_heads:
            raise ValueError("num_qo_heads must be a multiple of num_kv_heads")
        self._num_qo_heads, self._num_kv_heads = num_qo_heads, num_kv_heads
        self._head_dim_qk, self._head_dim_vo = head_dim_qk, head_dim_vo or head_dim_qk
        self._causal = bool(causal) and custom_mask is None  # MaskMode::kCustom overrides causal (reference prefill.cuh)
        causal = self._causal
        self._sm_scale = sm_scale if sm_scale is not None else 1.0 / math.sqrt(head_dim_qk)
        self._window_left = window_left
        self._logits_soft_cap = float(logits_soft_cap or 0.0)
        self._q_dtype = _canon_dtype(q_data_type)
        self._kv_dtype = _canon_dtype(kv_data_type) if kv_data_type is not None else self._q_dtype
        self._custom_mask = custom_mask
        qo_host = _host_i32(qo_indptr)
        self._qo_indptr_host = qo_host
        self._gen_cache_key = None  # device copies cached by the generic path belong to the previous plan
        self._kv_lens_host = kv_lens_host.to(torch.int32).contiguous()
        self._batch_size = qo_host.numel() - 1
        # ---- C++ LPT planner: (request, q-tile, q-head) units over the persistent grid ----
        num_ctas = device_sm_count(self.device if self.device.type == "cuda" else None)
        if self._cta_budget:
            num_ctas = max(1, min(num_ctas, int(self._cta_budget)))
        q_lens = (qo_host[1:] - qo_host[:-1]).tolist()
        max_work = sum((ql + _TILE_Q - 1) // _TILE_Q for ql in q_lens) * num_qo_heads
        pin32 = self._pin_int_workspace_buffer.view(torch.int32)
        need = max(max_work, 1) * _WORK_INTS + num_ctas + 1
        if need > pin32.numel():
            raise RuntimeError("int workspace too small for this batch")
        work = pin32[: max(max_work, 1) * _WORK_INTS]
        cta = pin32[max(max_work, 1) * _WORK_INTS : need]
        counts = torch.zeros(4, dtype=torch.int64)
        jit.load("planner").call(
            "prefill_plan", qo_host, self._kv_lens_host, self._kv_start_host, self._batch_size, num_qo_heads, _TILE_Q,
            _TILE_KV, 1 if causal else 0, int(window_left), num_ctas, work, max(max_work, 1), cta, counts,
        )
        self._num_work = int(counts[0])
        self._num_ctas = num_ctas
        dev32 = self._int_workspace_buffer.view(torch.int32)
        dev32[:need].copy_(pin32[:need], non_blocking=non_blocking and self.device.type == "cuda")
        self._work_info = dev32[: max(max_work, 1) * _WORK_INTS]
        self._cta_work_indptr = dev32[max(max_work, 1) * _WORK_INTS : need]
        self._mask_words = self._mask_bit_indptr = None
        if custom_mask is not None and self.device.type == "cuda":
            # packed little-endian bit stream (bit i = element i of the concatenated [q_len, kv_len] masks) as 32-bit words with two
            # spare words at the end (the kernel funnel-shifts 32 bits out of two neighbouring words), int64 bit offset per request
            from .attention.generic import pack_mask_bits

            by = pack_mask_bits(custom_mask.to(self.device).bool())
            pad = (-by.numel()) % 4 + 8
            self._mask_words = torch.cat([by, torch.zeros(pad, dtype=torch.uint8, device=self.device)]).view(torch.int32)
            bits = torch.zeros(self._batch_size + 1, dtype=torch.int64)
            bits[1:] = torch.cumsum((qo_host[1:] - qo_host[:-1]).long() * self._kv_lens_host.long(), 0)
            self._mask_bit_indptr = bits.to(self.device)
        self._planned = True

    def _run_reference(self, q, get_kv, out, lse, sm_scale, window_left):
        qo = self._qo_indptr_host
        outs = []
        for b in range(self._batch_size):
            qs, qe = int(qo[b]), int(qo[b + 1])
            if qe == qs:
                continue
            k, v = get_kv(b)
            mask = None
            if self._custom_mask is not None:
                off = sum(int(qo[i + 1] - qo[i]) * int(self._kv_lens_host[i]) for i in range(b))
                mask = self._custom_mask.flatten()[off : off + (qe - qs) * k.shape[0]].view(qe - qs, k.shape[0])
            if k.shape[0] == 0:
                o_b = torch.zeros(qe - qs, self._num_qo_heads, self._head_dim_vo, dtype=q.dtype, device=q.device)
                l_b = torch.full((qe - qs, self._num_qo_heads), float("-inf"), device=q.device)
            else:
                o_b, l_b = reference.attention_ref(q[qs:qe], k, v, self._causal and mask is None, sm_scale,
                                                   self._logits_soft_cap, window_left, custom_mask=mask,
                                                   alibi_slopes=self._alibi)
            out[qs:qe] = o_b.to(out.dtype)
            if lse is not None:
                lse[qs:qe] = l_b
        return outs

    def _launch_sm100(self, q, k, v, out, lse, sm_scale, window_left, paged, kv_indices, page_args, enable_pdl,
                      k_scale=None, v_scale=None, sinks=None):
        """Returns True when ``sinks`` were folded in by the kernel (tcgen05 path), False when the caller still has to."""
        shape_ok = ((self._head_dim_qk, self._head_dim_vo) in ((128, 128), (192, 128), (64, 64))
                    and q.dtype in (torch.float16, torch.bfloat16))
        if shape_ok and k.dtype in (torch.float8_e4m3fn, torch.float8_e5m2) and q.shape[0] >= 4 * self._batch_size:
            # fp8 KV with a compute-bound (prefill-sized) query: widen the KV that this call touches to the query dtype
            # once (1 B read + 2 B write per element, negligible next to the O(q * kv) attention work) and stay on the
            # tcgen05 kernel.  Scales were folded into sm_scale (k) and are applied to the output (v) by the caller.
            if paged:
                kv_indices_l = kv_indices.long()
                k = k.index_select(0, kv_indices_l).to(q.dtype)
                v = v.index_select(0, kv_indices_l).to(q.dtype)
                kv_indices = torch.arange(kv_indices.numel(), dtype=torch.int32, device=q.device)
                from .utils import paged_kv_strides as _pks

                sp, sn, sh, page_size, _, _ = _pks(k, self._kv_layout)
                page_args = (page_size, k.shape[0], sp, sn, sh, 1 if self._kv_layout == "HND" else 0)
            else:
                k, v = k.to(q.dtype), v.to(q.dtype)
                page_args = (1, k.shape[0], k.stride(0), k.stride(0), k.stride(1), 0, v.stride(0), v.stride(0), v.stride(1))
        fast = shape_ok and k.dtype == q.dtype and v.dtype == q.dtype
        if not fast:
            if self._variant_mod is not None:
                raise NotImplementedError("user attention variants run on the tcgen05 prefill kernel: f16 / bf16, head_dim 64 / 128 (192 qk)")
            self._launch_generic(q, k, v, out, lse, sm_scale, window_left, paged, kv_indices, page_args, enable_pdl, k_scale, v_scale)
            return False
        page_size, num_pages_total, sp, sn, sh, hnd = page_args[:6]
        vsp, vsn, vsh = page_args[6:9] if len(page_args) >= 9 else (sp, sn, sh)
        var_ptrs = var_scalars = None
        if self._variant_mod is not None:
            extra = list(getattr(self, "_variant_args", ()))
            nt, ns = len(self._variant_tensor_names), len(self._variant_scalar_names)
            if len(extra) != nt + ns:
                raise ValueError(f"variant expects {nt} tensors {self._variant_tensor_names} + {ns} scalars {self._variant_scalar_names} "
                                 f"after the q / kv arguments of run(), got {len(extra)}")
            self._variant_keepalive = [t.contiguous() for t in extra[:nt]]
            var_ptrs = torch.tensor([t.data_ptr() for t in self._variant_keepalive] + [0] * (8 - nt), dtype=torch.int64)
            var_scalars = torch.tensor([float(x) for x in extra[nt:]] + [0.0] * (8 - ns), dtype=torch.float64)
        (self._variant_mod or jit.load("prefill_sm100")).call(
            "prefill_run", q, k, v, out, lse, kv_indices, self._kv_page_indptr_dev if paged else None,
            self._work_info, self._cta_work_indptr, self._num_ctas, q.shape[0], self._num_qo_heads, self._num_kv_heads,
            self._head_dim_qk, 1 if paged else 0, page_size, num_pages_total, sp, sn, sh, vsp, vsn, vsh, hnd, q.stride(0),
            q.stride(1), out.stride(0), out.stride(1), float(sm_scale), float(self._logits_soft_cap), int(window_left),
            1 if self._causal else 0, self._mask_words, self._mask_bit_indptr, self._alibi,
            1 if self._variant_mod is not None else 0, var_ptrs, var_scalars,
            sinks.float().contiguous() if sinks is not None else None,
            dtype_code(q.dtype), 1 if (enable_pdl is None or enable_pdl) else 0, stream_ptr(q),
        )
        return sinks is not None

    def _check_plan_extras(self, q_data_type, o_data_type, prefix_len_ptr, token_pos_in_items_ptr, max_item_len_ptr) -> None:
        """Multi-item scoring masks are not implemented: refuse them at plan() time instead of computing plain attention.  An output
        dtype different from the query's is a conversion of the kernel's output (``_cast_output``)."""
        from .utils import reject_unsupported

        reject_unsupported("plan", prefix_len_ptr=prefix_len_ptr, token_pos_in_items_ptr=token_pos_in_items_ptr, max_item_len_ptr=max_item_len_ptr)
        self._o_cast = None
        if o_data_type is not None and _canon_dtype(o_data_type) != _canon_dtype(q_data_type):
            self._o_cast = _canon_dtype(o_data_type)

    def _cast_output(self, run, out):
        """run() with ``o_data_type != q_data_type``: the kernel writes the query dtype; the result is converted (into ``out`` when the
        caller passed a buffer of the planned output dtype)."""
        cast, self._o_cast = self._o_cast, None
        try:
            res = run()
        finally:
            self._o_cast = cast
        o = res[0] if isinstance(res, tuple) else res
        o = out.copy_(o) if out is not None else o.to(cast)
        return (o, *res[1:]) if isinstance(res, tuple) else o

    def _set_pos_encoding(self, pos_encoding_mode: str, num_qo_heads: int, rope_scale=None, rope_theta=None) -> None:
        """ALiBi is a logits transform of the softmax pass (slopes per head); ROPE_LLAMA rotates q and the touched keys before the
        kernel runs in its plain mode (attention/rope_on_the_fly.py)."""
        check_pos_encoding_mode(pos_encoding_mode)
        self._rope = None                                   # ROPE_LLAMA: (scale, theta); served by attention/rope_on_the_fly.py
        self._alibi = None
        if pos_encoding_mode == "ALIBI":
            from .utils import get_alibi_slopes

            self._alibi = get_alibi_slopes(num_qo_heads, self.device).float().contiguous()
        if pos_encoding_mode == "ROPE_LLAMA":
            from .attention.rope_on_the_fly import rope_params

            self._rope = rope_params(rope_scale, rope_theta)

    def _launch_generic(self, q, k, v, out, lse, sm_scale, window_left, paged, kv_indices, page_args, enable_pdl, k_scale, v_scale):
        """Catch-all CUDA-core kernel: other head dims, fp8 KV, custom masks."""
        from .attention import generic as _g

        if not _g.supported(q, k, self._head_dim_qk, self._head_dim_vo):
            raise NotImplementedError(f"attention: unsupported configuration (dtype {q.dtype}/{k.dtype}, "
                                      f"head_dim {self._head_dim_qk}/{self._head_dim_vo})")
        dev = q.device
        if getattr(self, "_gen_cache_key", None) != id(self._qo_indptr_host):
            self._gen_qo = self._qo_indptr_host.to(dev)
            if paged:
                self._gen_kv = self._kv_page_indptr_dev
                self._gen_last = self._kv_last_host.to(dev)
            else:
                self._gen_kv = None
                self._gen_last = None
            self._gen_mask = None
            if self._custom_mask is not None:
                self._gen_mask = _g.pack_mask_bits(self._custom_mask.to(dev))
                ql = (self._qo_indptr_host[1:] - self._qo_indptr_host[:-1]).long()
                bits = torch.zeros(self._batch_size + 1, dtype=torch.int64)
                bits[1:] = torch.cumsum(ql * self._kv_lens_host.long(), 0)
                self._gen_mask_indptr = bits.to(torch.int32).to(dev)
            self._gen_cache_key = id(self._qo_indptr_host)
        causal = self._causal and self._custom_mask is None
        if paged:
            page_size, _, sp, sn, sh, _ = page_args[:6]
            _g.run(q, k, v, out, lse, self._gen_qo, self._gen_kv, kv_indices, self._gen_last, page_size, (sp, sn, sh), (sp, sn, sh),
                   self._num_kv_heads, causal, window_left, sm_scale, self._logits_soft_cap, self._gen_mask,
                   getattr(self, "_gen_mask_indptr", None), self._alibi, 1.0, 1.0, enable_pdl is None or enable_pdl)
        else:
            # ragged: requests may be non-contiguous in k, so each request is one launch over its [start, start+len) slice
            for b in range(self._batch_size):
                qs, qe = int(self._qo_indptr_host[b]), int(self._qo_indptr_host[b + 1])
                if qe == qs:
                    continue
                st, ln = int(self._kv_start_host[b]), int(self._kv_lens_host[b])
                qi = torch.tensor([0, qe - qs], dtype=torch.int32, device=dev)
                ki = torch.tensor([0, ln], dtype=torch.int32, device=dev)
                pm = mi = None
                if self._gen_mask is not None:
                    off = int(self._gen_mask_indptr[b])
                    pm = _g.pack_mask_bits(self._custom_mask.flatten()[off: off + (qe - qs) * ln].to(dev))
                    mi = torch.tensor([0, (qe - qs) * ln], dtype=torch.int32, device=dev)
                _g.run(q[qs:qe], k[st:st + ln], v[st:st + ln], out[qs:qe], lse[qs:qe] if lse is not None else None, qi, ki, None,
                       None, 1, (0, k.stride(0), k.stride(1)), (0, v.stride(0), v.stride(1)), self._num_kv_heads, causal,
                       window_left, sm_scale, self._logits_soft_cap, pm, mi, self._alibi, 1.0, 1.0, enable_pdl is None or enable_pdl)

    def end_forward(self) -> None:
        pass


class BatchPrefillWithRaggedKVCacheWrapper(_BatchPrefillBase):
    """Batch prefill/append attention where K/V are ragged tensors ``[nnz_kv, Hkv, D]``."""

    def __init__(self, float_workspace_buffer, kv_layout: str = "NHD", use_cuda_graph: bool = False,
                 qo_indptr_buf=None, kv_indptr_buf=None, custom_mask_buf=None, mask_indptr_buf=None,
                 backend: str = "auto", jit_args=None, jit_kwargs=None):
        super().__init__(float_workspace_buffer, kv_layout, use_cuda_graph, qo_indptr_buf=qo_indptr_buf,
                         kv_indptr_buf=kv_indptr_buf, custom_mask_buf=custom_mask_buf, mask_indptr_buf=mask_indptr_buf,
                         backend=backend, jit_args=jit_args, jit_kwargs=jit_kwargs)

    def plan(self, qo_indptr, kv_indptr, num_qo_heads, num_kv_heads, head_dim_qk, head_dim_vo=None, custom_mask=None,
             packed_custom_mask=None, causal=False, pos_encoding_mode="NONE", use_fp16_qk_reduction=False,
             window_left=-1, logits_soft_cap=None, sm_scale=None, rope_scale=None, rope_theta=None,
             q_data_type="float16", kv_data_type=None, o_data_type=None, non_blocking=True, prefix_len_ptr=None,
             token_pos_in_items_ptr=None, token_pos_in_items_len=0, max_item_len_ptr=None, fixed_split_size=None,
             disable_split_kv=False, seq_lens=None, seq_lens_q=None, max_token_per_sequence=None, max_sequence_kv=None,
             v_indptr=None, o_indptr=None) -> None:
        """``seq_lens`` / ``seq_lens_q`` / ``max_token_per_sequence`` / ``max_sequence_kv`` restate what the indptr arrays say (the cuDNN
        backend of the reference wants them); separate element offsets for V and O (``v_indptr`` / ``o_indptr``) are refused."""
        from .utils import reject_unsupported

        reject_unsupported("BatchPrefillWithRaggedKVCacheWrapper.plan", v_indptr=v_indptr, o_indptr=o_indptr)
        remember_plan(self, locals())
        self._set_pos_encoding(pos_encoding_mode, num_qo_heads, rope_scale, rope_theta)
        self._check_plan_extras(q_data_type, o_data_type, prefix_len_ptr, token_pos_in_items_ptr, max_item_len_ptr)
        kv_host = kv_indptr.to("cpu", torch.int32)
        self._kv_start_host = kv_host[:-1].contiguous()
        self._kv_indptr_ragged_host = kv_host
        if packed_custom_mask is not None and custom_mask is None:
            qo_h = qo_indptr.to("cpu")
            custom_mask = _unpack_segmented(packed_custom_mask, ((qo_h[1:] - qo_h[:-1]) * (kv_host[1:] - kv_host[:-1])).tolist())
        self._plan_common(qo_indptr, kv_host[1:] - kv_host[:-1], num_qo_heads, num_kv_heads, head_dim_qk, head_dim_vo,
                          causal, sm_scale, window_left, logits_soft_cap, q_data_type, kv_data_type, custom_mask,
                          non_blocking)

    begin_forward = plan

    def run(self, q, k, v, *args, q_scale=None, k_scale=None, v_scale=None, o_scale=None, out=None, lse=None, return_lse=False,
            enable_pdl=None, window_left=None, kv_cache_sf=None):
        """``o_scale``: calibration scale of the output - the result is multiplied by ``v_scale / o_scale``.  NVFP4 K / V
        (``kv_cache_sf``) is refused."""
        if not self._planned:
            raise RuntimeError("plan() must be called before run()")
        if kv_cache_sf is not None:
            raise NotImplementedError("BatchPrefillWithRaggedKVCacheWrapper.run: NVFP4 K / V (kv_cache_sf) is not implemented")
        if o_scale is not None:
            v_scale = (1.0 if v_scale is None else float(v_scale)) / float(o_scale)
        if getattr(self, "_o_cast", None) is not None:
            return self._cast_output(lambda: self.run(q, k, v, *args, q_scale=q_scale, k_scale=k_scale, v_scale=v_scale, lse=lse,
                                                      return_lse=return_lse, enable_pdl=enable_pdl, window_left=window_left), out)
        self._variant_args = args
        if self._kv_layout == "HND":
            k, v = k.transpose(0, 1), v.transpose(0, 1)
        if self._rope is not None:                       # ROPE_LLAMA: rotate q and k, the kernel below runs in its plain mode
            from .attention.rope_on_the_fly import query_positions, ragged_key_positions, rotate_rows

            q = rotate_rows(q, query_positions(self._qo_indptr_host, self._kv_lens_host), *self._rope)
            k = rotate_rows(k.contiguous(), ragged_key_positions(self._kv_indptr_ragged_host), *self._rope)
        sm_scale = self._sm_scale * (q_scale or 1.0) * (k_scale or 1.0)
        window_left = self._window_left if window_left is None else window_left
        if out is None:
            out = torch.empty(q.shape[0], self._num_qo_heads, self._head_dim_vo, dtype=q.dtype, device=q.device)
        if return_lse and lse is None:
            lse = torch.empty(q.shape[0], self._num_qo_heads, dtype=torch.float32, device=q.device)
        if not q.is_cuda:
            st = self._kv_start_host
            self._run_reference(q, lambda b: (k[int(st[b]) : int(st[b]) + int(self._kv_lens_host[b])],
                                              v[int(st[b]) : int(st[b]) + int(self._kv_lens_host[b])]),
                                out, lse if return_lse else None, sm_scale, window_left)
        else:
            if k.stride(-1) != 1 or v.stride(-1) != 1:
                k, v = k.contiguous(), v.contiguous()
            page_args = (1, k.shape[0], k.stride(0), k.stride(0), k.stride(1), 0, v.stride(0), v.stride(0), v.stride(1))
            self._launch_sm100(q, k, v, out, lse if return_lse else None, sm_scale, window_left, False, None, page_args,
                               enable_pdl)
        if v_scale is not None:
            out.copy_((out.float() * v_scale).to(out.dtype))
        return (out, lse) if return_lse else out

    def forward(self, q, k, v, causal=False, pos_encoding_mode="NONE", use_fp16_qk_reduction=False, window_left=-1, logits_soft_cap=None,
                sm_scale=None, rope_scale=None, rope_theta=None):
        """Deprecated (use :meth:`run`): the attention parameters given here replace the planned ones, defaults included."""
        legacy_forward_replan(self, causal=causal, pos_encoding_mode=pos_encoding_mode, window_left=window_left, logits_soft_cap=logits_soft_cap,
                              sm_scale=sm_scale, rope_scale=rope_scale, rope_theta=rope_theta)
        return self.run(q, k, v)

    def forward_return_lse(self, q, k, v, causal=False, pos_encoding_mode="NONE", use_fp16_qk_reduction=False, window_left=-1,
                           logits_soft_cap=None, sm_scale=None, rope_scale=None, rope_theta=None):
        """Deprecated (use :meth:`run_return_lse`)."""
        legacy_forward_replan(self, causal=causal, pos_encoding_mode=pos_encoding_mode, window_left=window_left, logits_soft_cap=logits_soft_cap,
                              sm_scale=sm_scale, rope_scale=rope_scale, rope_theta=rope_theta)
        return self.run(q, k, v, return_lse=True)


class BatchPrefillWithPagedKVCacheWrapper(_BatchPrefillBase):
    """Batch prefill/append attention over a paged KV cache."""

    _paged = True

    def __init__(self, float_workspace_buffer, kv_layout: str = "NHD", use_cuda_graph: bool = False,
                 qo_indptr_buf=None, paged_kv_indptr_buf=None, paged_kv_indices_buf=None,
                 paged_kv_last_page_len_buf=None, custom_mask_buf=None, mask_indptr_buf=None, backend: str = "auto",
                 jit_args=None, jit_kwargs=None):
        super().__init__(float_workspace_buffer, kv_layout, use_cuda_graph, qo_indptr_buf=qo_indptr_buf,
                         paged_kv_indptr_buf=paged_kv_indptr_buf, paged_kv_indices_buf=paged_kv_indices_buf,
                         paged_kv_last_page_len_buf=paged_kv_last_page_len_buf, custom_mask_buf=custom_mask_buf,
                         mask_indptr_buf=mask_indptr_buf, backend=backend, jit_args=jit_args, jit_kwargs=jit_kwargs)

    def plan(self, qo_indptr, paged_kv_indptr, paged_kv_indices, paged_kv_last_page_len, num_qo_heads, num_kv_heads,
             head_dim_qk, page_size, head_dim_vo=None, custom_mask=None, packed_custom_mask=None, causal=False,
             pos_encoding_mode="NONE", use_fp16_qk_reduction=False, sm_scale=None, window_left=-1,
             logits_soft_cap=None, rope_scale=None, rope_theta=None, q_data_type="float16", kv_data_type=None,
             o_data_type=None, non_blocking=True, prefix_len_ptr=None, token_pos_in_items_ptr=None,
             token_pos_in_items_len=0, max_item_len_ptr=None, seq_lens=None, seq_lens_q=None, block_tables=None,
             max_token_per_sequence=None, max_sequence_kv=None, fixed_split_size=None, disable_split_kv=False) -> None:
        remember_plan(self, locals())
        self._set_pos_encoding(pos_encoding_mode, num_qo_heads, rope_scale, rope_theta)
        self._check_plan_extras(q_data_type, o_data_type, prefix_len_ptr, token_pos_in_items_ptr, max_item_len_ptr)
        self._page_size = page_size
        indptr_host = _host_i32(paged_kv_indptr)
        last_host = _host_i32(paged_kv_last_page_len)
        n_pages = indptr_host[1:] - indptr_host[:-1]
        kv_lens = torch.clamp(n_pages - 1, min=0) * page_size + torch.where(n_pages > 0, last_host, 0)
        self._kv_indptr_host, self._kv_last_host = indptr_host, last_host
        self._kv_start_host = indptr_host[:-1].contiguous()  # page-list start per request
        self._kv_indices = paged_kv_indices.to(self.device, torch.int32, non_blocking=non_blocking)
        self._kv_page_indptr_dev = indptr_host.to(self.device, non_blocking=non_blocking)
        if packed_custom_mask is not None and custom_mask is None:
            qo_h = qo_indptr.to("cpu")
            custom_mask = _unpack_segmented(packed_custom_mask, ((qo_h[1:] - qo_h[:-1]) * kv_lens).tolist())
        self._plan_common(qo_indptr, kv_lens, num_qo_heads, num_kv_heads, head_dim_qk, head_dim_vo, causal, sm_scale,
                          window_left, logits_soft_cap, q_data_type, kv_data_type, custom_mask, non_blocking)

    begin_forward = plan

    def run(self, q, paged_kv_cache, *args, q_scale=None, k_scale=None, v_scale=None, out=None, lse=None,
            return_lse=False, enable_pdl=None, window_left=None, sinks=None, kv_cache_sf=None,
            skip_softmax_threshold_scale_factor=None):
        """``skip_softmax_threshold_scale_factor`` (an approximation knob of the reference's trtllm-gen backend: KV tiles whose
        logits fall below a threshold skip the softmax) is accepted and not used - every tile is computed exactly.  NVFP4 KV
        pages (``kv_cache_sf``) are refused here (the decode wrapper reads them)."""
        if not self._planned:
            raise RuntimeError("plan() must be called before run()")
        if kv_cache_sf is not None:
            raise NotImplementedError("BatchPrefillWithPagedKVCacheWrapper.run: NVFP4 KV pages (kv_cache_sf) are not implemented; "
                                      "BatchDecodeWithPagedKVCacheWrapper reads them")
        if getattr(self, "_o_cast", None) is not None:
            return self._cast_output(lambda: self.run(q, paged_kv_cache, *args, q_scale=q_scale, k_scale=k_scale, v_scale=v_scale, lse=lse,
                                                      return_lse=return_lse, enable_pdl=enable_pdl, window_left=window_left, sinks=sinks), out)
        self._variant_args = args
        user_return_lse = return_lse
        return_lse = return_lse or sinks is not None
        k_cache, v_cache = unpack_paged_kv_cache(paged_kv_cache, self._kv_layout)
        if self._rope is not None:                       # ROPE_LLAMA: rotate q and the batch's key pages (scratch copy of the cache)
            from .attention.rope_on_the_fly import query_positions, rotate_rows, rotated_paged_keys

            q = rotate_rows(q, query_positions(self._qo_indptr_host, self._kv_lens_host), *self._rope)
            k_cache = rotated_paged_keys(k_cache, self._kv_indices, self._kv_indptr_host, self._kv_layout, *self._rope)
        sm_scale = self._sm_scale * (q_scale or 1.0) * (k_scale or 1.0)
        window_left = self._window_left if window_left is None else window_left
        if out is None:
            out = torch.empty(q.shape[0], self._num_qo_heads, self._head_dim_vo, dtype=q.dtype, device=q.device)
        if return_lse and lse is None:
            lse = torch.empty(q.shape[0], self._num_qo_heads, dtype=torch.float32, device=q.device)
        if not q.is_cuda:
            self._run_reference(
                q, lambda b: reference.gather_paged_kv(k_cache, v_cache, self._kv_indices.cpu(), self._kv_indptr_host,
                                                       self._kv_last_host, b, self._kv_layout),
                out, lse if return_lse else None, sm_scale, window_left)
        else:
            sp, sn, sh, page_size, hkv, d = paged_kv_strides(k_cache, self._kv_layout)
            if paged_kv_strides(v_cache, self._kv_layout)[:3] != (sp, sn, sh):
                raise ValueError("k_cache and v_cache must share strides")
            page_args = (page_size, k_cache.shape[0], sp, sn, sh, 1 if self._kv_layout == "HND" else 0)
            folded = self._launch_sm100(q, k_cache, v_cache, out, lse if return_lse else None, sm_scale, window_left, True,
                                        self._kv_indices, page_args, enable_pdl, sinks=sinks)
            if folded:
                sinks = None  # exp(sink) joined the softmax denominator inside the kernel
        if sinks is not None:
            from .attention._core import apply_attention_sink

            o2, l2 = apply_attention_sink(out, lse, sinks)
            out.copy_(o2)
            lse.copy_(l2)
        if v_scale is not None:
            out.copy_((out.float() * v_scale).to(out.dtype))
        return (out, lse) if user_return_lse else out

    def forward(self, q, paged_kv_cache, causal=False, pos_encoding_mode="NONE", use_fp16_qk_reduction=False, k_scale=None, v_scale=None,
                window_left=-1, logits_soft_cap=None, sm_scale=None, rope_scale=None, rope_theta=None):
        """Deprecated (use :meth:`run`): the attention parameters given here replace the planned ones, defaults included."""
        legacy_forward_replan(self, causal=causal, pos_encoding_mode=pos_encoding_mode, window_left=window_left, logits_soft_cap=logits_soft_cap,
                              sm_scale=sm_scale, rope_scale=rope_scale, rope_theta=rope_theta)
        return self.run(q, paged_kv_cache, k_scale=k_scale, v_scale=v_scale)

    def forward_return_lse(self, q, paged_kv_cache, causal=False, pos_encoding_mode="NONE", use_fp16_qk_reduction=False, k_scale=None,
                           v_scale=None, window_left=-1, logits_soft_cap=None, sm_scale=None, rope_scale=None, rope_theta=None):
        """Deprecated (use :meth:`run_return_lse`)."""
        legacy_forward_replan(self, causal=causal, pos_encoding_mode=pos_encoding_mode, window_left=window_left, logits_soft_cap=logits_soft_cap,
                              sm_scale=sm_scale, rope_scale=rope_scale, rope_theta=rope_theta)
        return self.run(q, paged_kv_cache, k_scale=k_scale, v_scale=v_scale, return_lse=True)


# ------------------------------------------------------------------------------------------------
# Function-style context (prefill) APIs.  Parity: reference flashinfer/prefill.py:3557-4435
# (fmha_varlen, trtllm_ragged_attention_deepseek, trtllm_batch_context_with_kv_cache, trtllm_fmha_v2_prefill,
# cudnn_batch_prefill_with_kv_cache).  One kernel family on B200, so they share the wrappers above.
# ------------------------------------------------------------------------------------------------
def fmha_varlen(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, qo_segment_offsets: torch.Tensor,
                kv_segment_offsets: torch.Tensor, plan_info=None, max_qo_len: Optional[int] = None, out=None, lse=None,
                causal: bool = False, sm_scale: Optional[float] = None, q_scale: Optional[float] = None,
                k_scale: Optional[float] = None, v_scale: Optional[float] = None, o_scale: Optional[float] = None,
                return_lse: bool = False):
    """Ragged variable-length FMHA (q/k/v ``[nnz, H, D]`` + segment offsets).  ``q_scale * k_scale`` folds into the softmax scale;
    the output is multiplied by ``v_scale / o_scale`` (reference prefill.py :3600)."""
    ws = torch.empty(16 << 20, dtype=torch.uint8, device=q.device)
    w = BatchPrefillWithRaggedKVCacheWrapper(ws)
    w.plan(qo_segment_offsets, kv_segment_offsets, q.shape[1], k.shape[1], q.shape[2], head_dim_vo=v.shape[2],
           causal=causal, sm_scale=sm_scale, q_data_type=q.dtype)
    return w.run(q, k, v, out=out, lse=lse, return_lse=return_lse, q_scale=q_scale, k_scale=k_scale, v_scale=v_scale, o_scale=o_scale)


def fmha_varlen_plan(module, qo_segment_offsets, kv_segment_offsets, num_qo_heads, causal):
    """The reference runs a device-side plan kernel here; our C++ planner is invoked inside ``fmha_varlen``."""
    return None


def trtllm_ragged_attention_deepseek(query, key, value, workspace_buffer, seq_lens, max_q_len, max_kv_len, bmm1_scale,
                                     bmm2_scale, o_sf_scale, batch_size, window_left, cum_seq_lens_q, cum_seq_lens_kv,
                                     enable_pdl=False, is_causal=True, return_lse=False, attention_sinks=None,
                                     skip_softmax_threshold_scale_factor=None, out=None, lse=None, sage_attn_sfs=(None, None, None, None),
                                     num_elts_per_sage_attn_blk=(0, 0, 0, 0), backend: str = "trtllm-gen"):
    """``backend`` names the reference's kernel family (one native kernel here); SageAttention block scales are refused."""
    if any(t is not None for t in sage_attn_sfs) or any(int(n) != 0 for n in num_elts_per_sage_attn_blk):
        raise NotImplementedError("trtllm_ragged_attention_deepseek: SageAttention block scales (sage_attn_sfs) are not implemented")
    if o_sf_scale is not None and float(o_sf_scale) > 0:
        raise NotImplementedError("trtllm_ragged_attention_deepseek: NVFP4 output (o_sf_scale > 0) is not implemented")
    w = BatchPrefillWithRaggedKVCacheWrapper(workspace_buffer)
    w.plan(cum_seq_lens_q, cum_seq_lens_kv, query.shape[1], key.shape[1], query.shape[2], head_dim_vo=value.shape[2],
           causal=is_causal, sm_scale=float(bmm1_scale), window_left=window_left, q_data_type=query.dtype)
    vs = float(bmm2_scale) if float(bmm2_scale) != 1.0 else None
    if attention_sinks is None:
        return w.run(query, key, value, out=out, lse=lse, return_lse=return_lse, v_scale=vs)
    # per-head sink logits join the softmax denominator through the (out, lse) state
    from .attention._core import apply_attention_sink

    o, l = w.run(query, key, value, lse=lse, return_lse=True, v_scale=vs)
    o2, l2 = apply_attention_sink(o, l, attention_sinks.to(o.device))
    if out is not None:
        out.copy_(o2)
        o2 = out
    if lse is not None:
        lse.copy_(l2)
        l2 = lse
    return (o2, l2) if return_lse else o2


def trtllm_batch_context_with_kv_cache(query, kv_cache, workspace_buffer, block_tables, seq_lens, max_q_len, max_kv_len,
                                       bmm1_scale, bmm2_scale, batch_size, cum_seq_lens_q, cum_seq_lens_kv,
                                       window_left: int = -1, out=None, out_dtype=None, o_sf_scale=None,
                                       o_sf_vec_size=None, kv_layout: str = "HND", enable_pdl=None, sinks=None,
                                       kv_cache_sf=None, skip_softmax_threshold_scale_factor=None,
                                       uses_shared_paged_kv_idx: bool = True, causal: bool = True, lse=None,
                                       return_lse: bool = False):
    """Paged context attention with a block-table interface; ``causal=False`` = dense / bidirectional attention (no sliding
    window then, like the reference :4155)."""
    if not causal and window_left >= 0:
        raise ValueError("Sliding-window non-causal attention is not supported for the paged context path")
    from .decode import _block_tables_to_indices
    from .utils import reject_unsupported

    # NVFP4 output / NVFP4 KV scale factors are not implemented on the context path; skip-softmax is a speed hint (exact here)
    reject_unsupported("trtllm_batch_context_with_kv_cache", o_sf_scale=o_sf_scale, o_sf_vec_size=o_sf_vec_size, kv_cache_sf=kv_cache_sf,
                       uses_shared_paged_kv_idx=(uses_shared_paged_kv_idx, True))

    k_cache, v_cache = unpack_paged_kv_cache(kv_cache, kv_layout)
    _, _, _, page_size, hkv, d = paged_kv_strides(k_cache, kv_layout)
    indptr, indices, last = _block_tables_to_indices(block_tables, seq_lens, page_size)
    w = BatchPrefillWithPagedKVCacheWrapper(workspace_buffer, kv_layout)
    w.plan(cum_seq_lens_q, indptr, indices, last, query.shape[1], hkv, d, page_size, causal=bool(causal),
           sm_scale=float(bmm1_scale), window_left=window_left, q_data_type=query.dtype)
    res = w.run(query, (k_cache, v_cache), out=out if (out is None or out.dtype == query.dtype) else None, lse=lse,
                return_lse=return_lse, sinks=sinks, v_scale=float(bmm2_scale) if float(bmm2_scale) != 1.0 else None)
    want = out.dtype if out is not None else out_dtype
    if want is None or want == query.dtype:
        return res
    o = (res[0] if return_lse else res).to(want)          # output dtype other than the query's (e.g. fp8 out): converted after the kernel
    if out is not None:
        out.copy_(o)
        o = out
    return (o, res[1]) if return_lse else o


def fmha_v2_prefill_deepseek(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, out: torch.Tensor, num_heads: int,
                             head_dim: int, seq_len: int, scale_softmax: float, scale_bmm1: Optional[float] = None,
                             scale_bmm2: Optional[float] = None, return_lse: bool = False, lse: Optional[torch.Tensor] = None):
    """DeepSeek-R1 context attention (reference prefill.py:4345, an sm_120 fmha_v2 kernel there): ``query / key [B, S, H, 192]``,
    ``value [B, S, H, 128]``, causal.  Runs the tcgen05 192 / 128 ragged prefill kernel; ``scale_softmax`` (0 = default
    ``1 / sqrt(192)``) and the bmm scales fold into the softmax scale / output."""
    B, S, H, dqk = query.shape
    dvo = value.shape[-1]
    sm = float(scale_softmax) if scale_softmax else dqk ** -0.5
    sm *= float(scale_bmm1) if scale_bmm1 is not None else 1.0
    indptr = torch.arange(0, (B + 1) * S, S, dtype=torch.int32)
    w = BatchPrefillWithRaggedKVCacheWrapper(torch.empty(64 << 20, dtype=torch.uint8, device=query.device))
    w.plan(indptr, indptr, H, H, dqk, head_dim_vo=dvo, causal=True, sm_scale=sm, q_data_type=query.dtype)
    res = w.run(query.reshape(B * S, H, dqk), key.reshape(B * S, H, dqk), value.reshape(B * S, H, dvo), return_lse=return_lse)
    o, l = res if return_lse else (res, None)
    if scale_bmm2 is not None and float(scale_bmm2) != 1.0:
        o = (o.float() * float(scale_bmm2)).to(o.dtype)
    out.copy_(o.view(B, S, H, dvo).to(out.dtype))
    if return_lse:
        l = l.view(B, S, H)
        if lse is not None:
            lse.copy_(l)
            l = lse
        return out, l
    return out


def trtllm_fmha_v2_prefill(*args, **kwargs):
    """sm90/sm120-only TRT-LLM fmha_v2 path in the reference (jit/attention/modules.py:2010); B200 uses fmha_varlen."""
    raise NotImplementedError("fmha_v2 is an sm90/sm120 path; use fmha_varlen / the prefill wrappers on B200")


def cudnn_batch_prefill_with_kv_cache(q, k_cache, v_cache, scale, workspace_buffer, *, max_token_per_sequence,
                                      max_sequence_kv, actual_seq_lens_q, actual_seq_lens_kv, block_tables=None,
                                      causal: bool = True, return_lse: bool = False, batch_offsets_q=None,
                                      batch_offsets_o=None, is_cuda_graph_compatible=False, out=None, lse=None, **kw):
    """cuDNN-style prefill signature (reference flashinfer/cudnn/prefill.py:563)."""
    from .decode import _block_tables_to_indices

    ql = actual_seq_lens_q.reshape(-1).to("cpu", torch.int64)
    qo = torch.zeros(ql.numel() + 1, dtype=torch.int32)
    qo[1:] = ql.cumsum(0)
    if block_tables is not None:
        _, _, _, page_size, hkv, d = paged_kv_strides(k_cache, "HND")
        indptr, indices, last = _block_tables_to_indices(block_tables, actual_seq_lens_kv.reshape(-1), page_size)
        w = BatchPrefillWithPagedKVCacheWrapper(workspace_buffer, "HND")
        w.plan(qo, indptr, indices, last, q.shape[1], hkv, d, page_size, causal=causal, sm_scale=scale,
               q_data_type=q.dtype)
        return w.run(q, (k_cache, v_cache), out=out, lse=lse, return_lse=return_lse)
    kl = actual_seq_lens_kv.reshape(-1).to("cpu", torch.int64)
    kvi = torch.zeros(kl.numel() + 1, dtype=torch.int32)
    kvi[1:] = kl.cumsum(0)
    w = BatchPrefillWithRaggedKVCacheWrapper(workspace_buffer)
    w.plan(qo, kvi, q.shape[1], k_cache.shape[1], q.shape[2], causal=causal, sm_scale=scale, q_data_type=q.dtype)
    return w.run(q, k_cache, v_cache, out=out, lse=lse, return_lse=return_lse)


from . import jit as _jit_acc  # noqa: E402

get_batch_prefill_module = _jit_acc.module_accessor("prefill_sm100")
get_batch_prefill_jit_module = _jit_acc.module_accessor("prefill_sm100")
get_single_prefill_module = _jit_acc.module_accessor("prefill_sm100")
get_customize_batch_prefill_module = _jit_acc.module_accessor("attention_generic")
get_fmha_module = _jit_acc.module_accessor("prefill_sm100")
get_trtllm_fmha_v2_module = _jit_acc.module_accessor("prefill_sm100")
get_trtllm_gen_prefill_module = _jit_acc.module_accessor("prefill_sm100")
get_trtllm_gen_fmha_module = _jit_acc.module_accessor("prefill_sm100")


def single_prefill_with_kv_cache_with_jit_module(jit_module, q, k, v, *args, kv_layout: str = "NHD", mask_mode: int = 0,
                                                 window_left: int = -1, return_lse: bool = False, **kwargs):
    """Reference prefill.py: single-request prefill through an explicitly supplied JIT module; ``mask_mode`` 1 = causal."""
    return single_prefill_with_kv_cache(q, k, v, causal=(mask_mode == 1), kv_layout=kv_layout, window_left=window_left,
                                        return_lse=return_lse, **kwargs)


def make_hashable_cache(func):
    """``functools.cache`` that tolerates unhashable arguments (lists / dicts are frozen into tuples) - the reference uses it
    to memoise JIT-module getters keyed by lists of extra tensor names."""
    import functools

    def freeze(x):
        if isinstance(x, (list, tuple)):
            return tuple(freeze(i) for i in x)
        if isinstance(x, dict):
            return tuple(sorted((k, freeze(v)) for k, v in x.items()))
        return x

    memo = {}

    @functools.wraps(func)
    def wrapper(*args, **kwargs):
        key = (freeze(args), freeze(kwargs))
        if key not in memo:
            memo[key] = func(*args, **kwargs)
        return memo[key]

    wrapper.cache_clear = memo.clear
    return wrapper
