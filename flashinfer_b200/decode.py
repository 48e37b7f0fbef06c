"""Decode attention: ``single_decode_with_kv_cache`` and ``BatchDecodeWithPagedKVCacheWrapper``.

API parity: reference flashinfer/decode.py:409-604 (single decode), :606-1603 (batch wrapper,
plan/run, CUDA-graph mode), :1605 (CUDAGraph wrapper), :2897 (fast_decode_plan).

B200-first implementation: one persistent tcgen05 kernel (csrc/attention/decode_sm100.cu) driven by
a C++ stream-K planner (csrc/runtime/planner.cpp).  There is no backend zoo: CUDA tensors always
run the sm_100a kernel (and fail loudly when a configuration is not specialised); CPU tensors run
the fp32 PyTorch oracle.
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

_SEG_INTS = 12
_MERGE_INTS = 8
_TILE_KV = 128
_MAX_Q_ROWS = 32


def _canon_dtype(dt):
    if isinstance(dt, str):
        return getattr(torch, dt)
    return dt


def single_decode_with_kv_cache(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    kv_layout: str = "NHD",
    pos_encoding_mode: str = "NONE",
    use_tensor_cores: bool = False,
    q_scale: Optional[float] = None,
    k_scale: Optional[float] = None,
    v_scale: Optional[float] = None,
    window_left: int = -1,
    logits_soft_cap: Optional[float] = None,
    sm_scale: Optional[float] = None,
    rope_scale: Optional[float] = None,
    rope_theta: Optional[float] = None,
    return_lse: bool = False,
):
    """Decode attention for one request: q [Hq, D], k/v [kv_len, Hkv, D] (NHD) or [Hkv, kv_len, D] (HND)."""
    check_kv_layout(kv_layout)
    check_pos_encoding_mode(pos_encoding_mode)
    if pos_encoding_mode == "ALIBI":                    # the bias pass lives in the prefill kernel: one query row there
        from .prefill import single_prefill_with_kv_cache

        scale = (sm_scale if sm_scale is not None else 1.0 / math.sqrt(q.shape[-1])) * (q_scale or 1.0) * (k_scale or 1.0)
        res = single_prefill_with_kv_cache(q.unsqueeze(0), k, v, causal=False, kv_layout=kv_layout, pos_encoding_mode="ALIBI", sm_scale=scale,
                                           window_left=window_left, logits_soft_cap=logits_soft_cap, return_lse=True)
        o = res[0][0] if v_scale is None else (res[0][0].float() * v_scale).to(q.dtype)
        return (o, res[1][0]) if return_lse else o
    head_dim = q.shape[-1]
    if sm_scale is None:
        sm_scale = 1.0 / math.sqrt(head_dim)
    if q_scale is not None:
        sm_scale *= q_scale
    if k_scale is not None:
        sm_scale *= k_scale
    kn = k if kv_layout == "NHD" else k.transpose(0, 1)
    vn = v if kv_layout == "NHD" else v.transpose(0, 1)
    if pos_encoding_mode == "ROPE_LLAMA":              # rotate q (the newest position) and k, then plain attention
        from .attention.rope_on_the_fly import rope_params, rotate_rows

        rs, rt = rope_params(rope_scale, rope_theta)
        q = rotate_rows(q.unsqueeze(0), torch.tensor([kn.shape[0] - 1]), rs, rt)[0]
        kn = rotate_rows(kn.contiguous(), torch.arange(kn.shape[0]), rs, rt)
    if not q.is_cuda:
        o, lse = reference.attention_ref(
            q.unsqueeze(0), kn, vn, False, sm_scale, logits_soft_cap or 0.0, window_left
        )
        o, lse = o[0], lse[0]
    else:
        # run through the paged kernel with the contiguous KV viewed as pages of 128 tokens
        kv_len = kn.shape[0]
        hkv = kn.shape[1]
        page = _TILE_KV
        n_pages = max(1, (kv_len + page - 1) // page)
        pad = n_pages * page - kv_len
        if pad:
            kn = torch.cat([kn, kn.new_zeros(pad, hkv, head_dim)], 0)
            vn = torch.cat([vn, vn.new_zeros(pad, hkv, head_dim)], 0)
        kc = kn.reshape(n_pages, page, hkv, head_dim)
        vc = vn.reshape(n_pages, page, hkv, head_dim)
        ws = torch.empty(8 * 1024 * 1024, dtype=torch.uint8, device=q.device)
        w = BatchDecodeWithPagedKVCacheWrapper(ws, "NHD")
        indptr = torch.tensor([0, n_pages], dtype=torch.int32)
        indices = torch.arange(n_pages, dtype=torch.int32)
        last = torch.tensor([kv_len - (n_pages - 1) * page], dtype=torch.int32)
        w.plan(
            indptr, indices, last, q.shape[0], hkv, head_dim, page,
            window_left=window_left, logits_soft_cap=logits_soft_cap, q_data_type=q.dtype, sm_scale=sm_scale,
        )
        o, lse = w.run(q.unsqueeze(0), (kc, vc), return_lse=True)
        o, lse = o[0], lse[0]
    if v_scale is not None:
        o = (o.float() * v_scale).to(o.dtype)
    return (o, lse) if return_lse else o


class BatchDecodeWithPagedKVCacheWrapper:
    """Batch decode over a paged KV cache (plan once per batch composition, run once per layer)."""

    def __init__(
        self,
        float_workspace_buffer: torch.Tensor,
        kv_layout: str = "NHD",
        use_cuda_graph: bool = False,
        use_tensor_cores: bool = True,
        paged_kv_indptr_buffer: Optional[torch.Tensor] = None,
        paged_kv_indices_buffer: Optional[torch.Tensor] = None,
        paged_kv_last_page_len_buffer: Optional[torch.Tensor] = None,
        backend: str = "auto",
        jit_args=None,
    ) -> None:
        check_kv_layout(kv_layout)
        self._kv_layout = kv_layout
        self._float_workspace_buffer = float_workspace_buffer
        self.device = float_workspace_buffer.device
        self._use_cuda_graph = use_cuda_graph
        self._int_workspace_buffer = torch.empty(8 * 1024 * 1024, dtype=torch.uint8, device=self.device)
        self._pin_int_workspace_buffer = torch.empty(
            8 * 1024 * 1024, dtype=torch.uint8, device="cpu", pin_memory=self.device.type == "cuda"
        )
        # in-kernel split-KV merge: one self-resetting arrival counter per partial-state slot
        self._merge_counters = torch.zeros(4096, dtype=torch.int32, device=self.device)
        if use_cuda_graph:
            if paged_kv_indptr_buffer is None or paged_kv_indices_buffer is None or paged_kv_last_page_len_buffer is None:
                raise ValueError("use_cuda_graph=True requires the indptr/indices/last_page_len buffers")
            self._fixed_batch_size = paged_kv_last_page_len_buffer.numel()
        self._paged_kv_indptr_buf = paged_kv_indptr_buffer
        self._paged_kv_indices_buf = paged_kv_indices_buffer
        self._paged_kv_last_page_len_buf = paged_kv_last_page_len_buffer
        self._planned = False
        self._backend = "sm100"
        self._cta_budget: Optional[int] = None  # POD: restrict the persistent grid to this many SMs

    @property
    def use_tensor_cores(self) -> bool:
        return True

    @property
    def is_cuda_graph_enabled(self) -> bool:
        return self._use_cuda_graph

    def reset_workspace_buffer(self, float_workspace_buffer: torch.Tensor, int_workspace_buffer: torch.Tensor) -> None:
        self._float_workspace_buffer = float_workspace_buffer
        self._int_workspace_buffer = int_workspace_buffer
        self._pin_int_workspace_buffer = torch.empty(
            int_workspace_buffer.numel(), dtype=torch.uint8, device="cpu", pin_memory=self.device.type == "cuda"
        )

    # ------------------------------------------------------------------ plan
    def plan(
        self,
        indptr: torch.Tensor,
        indices: torch.Tensor,
        last_page_len: torch.Tensor,
        num_qo_heads: int,
        num_kv_heads: int,
        head_dim: int,
        page_size: int,
        pos_encoding_mode: str = "NONE",
        window_left: int = -1,
        logits_soft_cap: Optional[float] = None,
        q_data_type: Union[str, torch.dtype] = "float16",
        kv_data_type: Optional[Union[str, torch.dtype]] = None,
        o_data_type: Optional[Union[str, torch.dtype]] = None,
        data_type: Optional[Union[str, torch.dtype]] = None,
        sm_scale: Optional[float] = None,
        rope_scale: Optional[float] = None,
        rope_theta: Optional[float] = None,
        non_blocking: bool = True,
        block_tables: Optional[torch.Tensor] = None,
        seq_lens: Optional[torch.Tensor] = None,
        fixed_split_size: Optional[int] = None,
        disable_split_kv: bool = False,
        qo_indptr: Optional[torch.Tensor] = None,
    ) -> None:
        """Host-side planning.  ``qo_indptr`` (extension) allows q_len>1 per request
        (speculative decode / small append) as long as ``q_len * group <= 32``."""
        remember_plan(self, locals())
        self._qlen_planned = 1
        check_pos_encoding_mode(pos_encoding_mode)
        self._alibi_delegate = None
        if pos_encoding_mode == "ALIBI":                # the bias pass lives in the prefill kernel: delegate with one query row per request
            from .prefill import BatchPrefillWithPagedKVCacheWrapper

            bsz = last_page_len.numel()
            self._alibi_delegate = BatchPrefillWithPagedKVCacheWrapper(self._float_workspace_buffer, self._kv_layout)
            self._alibi_delegate.plan(qo_indptr if qo_indptr is not None else torch.arange(bsz + 1, dtype=torch.int32), indptr, indices, last_page_len,
                                      num_qo_heads, num_kv_heads, head_dim, page_size, causal=qo_indptr is not None, pos_encoding_mode="ALIBI",
                                      sm_scale=sm_scale, window_left=window_left, logits_soft_cap=logits_soft_cap,
                                      q_data_type=data_type if data_type is not None else q_data_type, kv_data_type=kv_data_type)
            self._planned = True
            return
        self._rope = None                                   # ROPE_LLAMA: (scale, theta); served by attention/rope_on_the_fly.py
        if pos_encoding_mode == "ROPE_LLAMA":
            from .attention.rope_on_the_fly import rope_params

            self._rope = rope_params(rope_scale, rope_theta)
        if num_qo_heads % num_kv_heads != 0:
            raise ValueError("num_qo_heads must be a multiple of num_kv_heads")
        batch_size = last_page_len.numel()
        if indptr.numel() != batch_size + 1:
            raise ValueError("indptr must have batch_size + 1 entries")
        q_dt = _canon_dtype(data_type if data_type is not None else q_data_type)
        kv_dt = _canon_dtype(kv_data_type) if kv_data_type is not None else q_dt
        self._q_dtype, self._kv_dtype = q_dt, kv_dt
        self._o_dtype = _canon_dtype(o_data_type) if o_data_type is not None else q_dt
        self._num_qo_heads, self._num_kv_heads = num_qo_heads, num_kv_heads
        self._head_dim, self._page_size = head_dim, page_size
        self._window_left = window_left
        self._logits_soft_cap = float(logits_soft_cap or 0.0)
        self._sm_scale = sm_scale if sm_scale is not None else 1.0 / math.sqrt(head_dim)
        self._batch_size = batch_size

        indptr_host = _host_i32(indptr)
        last_host = _host_i32(last_page_len)
        n_pages = indptr_host[1:] - indptr_host[:-1]
        kv_lens_host = (torch.clamp(n_pages - 1, min=0) * page_size + torch.where(n_pages > 0, last_host, 0)).to(torch.int32)
        qo_host = _host_i32(qo_indptr) if qo_indptr is not None else None
        self._qo_indptr_host = qo_host
        self._kv_lens_host = kv_lens_host

        if self._use_cuda_graph:
            if batch_size != self._fixed_batch_size:
                raise ValueError("batch size must stay fixed under CUDA graphs")
            self._paged_kv_indptr_buf.copy_(indptr, non_blocking=non_blocking)
            self._paged_kv_indices_buf[: indices.numel()].copy_(indices, non_blocking=non_blocking)
            self._paged_kv_last_page_len_buf.copy_(last_page_len, non_blocking=non_blocking)
            self._kv_indices = self._paged_kv_indices_buf
        else:
            self._kv_indices = indices.to(self.device, torch.int32, non_blocking=non_blocking)
        self._kv_indptr_host = indptr_host
        self._kv_last_host = last_host
        self._gen_key = None  # device copies cached by the generic path belong to the previous plan

        # ---- C++ planner into the pinned buffer, then ONE H2D copy ----
        num_ctas = device_sm_count(self.device if self.device.type == "cuda" else None)
        if self._cta_budget:
            num_ctas = max(1, min(num_ctas, int(self._cta_budget)))
        group = num_qo_heads // num_kv_heads
        max_segs = batch_size * num_kv_heads + num_ctas + 1
        max_merge = num_ctas + 1
        pin32 = self._pin_int_workspace_buffer.view(torch.int32)
        need = max_segs * _SEG_INTS + (num_ctas + 1) + max_merge * _MERGE_INTS
        if need > pin32.numel():
            raise RuntimeError("int workspace too small for this batch")
        seg = pin32[: max_segs * _SEG_INTS]
        cta = pin32[max_segs * _SEG_INTS : max_segs * _SEG_INTS + num_ctas + 1]
        mrg = pin32[max_segs * _SEG_INTS + num_ctas + 1 : need]
        counts = torch.zeros(8, dtype=torch.int64)
        planner = jit.load("planner")
        planner.call(
            "decode_plan",
            indptr_host.contiguous(), kv_lens_host.contiguous(), qo_host, batch_size, num_kv_heads, group,
            page_size, num_ctas, 2 if not disable_split_kv else 1 << 30, seg, max_segs, cta, mrg, max_merge, counts,
        )
        nseg, nmerge, nslots, max_q_rows = (int(x) for x in counts[:4])
        self._num_merge = nmerge
        self._num_slots = nslots
        self._max_q_rows = max_q_rows
        self._num_ctas = num_ctas
        self._plan_counts = counts
        dev32 = self._int_workspace_buffer.view(torch.int32)
        if self.device.type == "cuda":
            dev32[:need].copy_(pin32[:need], non_blocking=non_blocking)
        else:
            dev32[:need].copy_(pin32[:need])
        self._seg_info = dev32[: max_segs * _SEG_INTS]
        self._cta_seg_indptr = dev32[max_segs * _SEG_INTS : max_segs * _SEG_INTS + num_ctas + 1]
        self._merge_items = dev32[max_segs * _SEG_INTS + num_ctas + 1 : need]
        nv = 1
        while nv < max_q_rows:
            nv *= 2
        self._rows_per_slot = nv
        slots_cap = max(nslots, 1) if not self._use_cuda_graph else 2 * num_ctas + 2
        fneed = slots_cap * nv * (head_dim + 1) * 4
        if fneed > self._float_workspace_buffer.numel() * self._float_workspace_buffer.element_size():
            raise RuntimeError("float workspace too small for split-KV partials")
        fws = self._float_workspace_buffer.view(torch.uint8)[: fneed].view(torch.float32)
        self._partial_o = fws[: slots_cap * nv * head_dim]
        self._partial_lse = fws[slots_cap * nv * head_dim :]
        self._planned = True

    begin_forward = plan

    # ------------------------------------------------------------------ run
    def run(
        self,
        q: torch.Tensor,
        paged_kv_cache: Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]],
        *args,
        q_scale: Optional[float] = None,
        k_scale: Optional[float] = None,
        v_scale: Optional[float] = None,
        out: Optional[torch.Tensor] = None,
        lse: Optional[torch.Tensor] = None,
        return_lse: bool = False,
        enable_pdl: Optional[bool] = None,
        window_left: Optional[int] = None,
        sinks: Optional[torch.Tensor] = None,
        kv_cache_sf=None,
        kv_prefetch: bool = False,
        q_len_per_req: Optional[int] = 1,
        skip_softmax_threshold_scale_factor: Optional[float] = None,
    ):
        """``q_len_per_req`` (reference decode.py :1285): query tokens per request, ``q [batch * q_len_per_req, H, D]`` - more than
        one (speculative verification, causal over the newest tokens) plans the multi-token form again from the remembered
        ``plan()`` arguments.  ``skip_softmax_threshold_scale_factor`` (an approximation knob of the reference's trtllm-gen backend)
        is accepted and not used: every KV tile is computed exactly.

        ``kv_prefetch=True`` (with PDL): promise that the kernel launched just before this one on the stream only APPENDS the
        newest ``q_len`` tokens of each request to the cache (an append / fused QKV+RoPE+append kernel): the TMA producers then
        stream all older KV tiles before the programmatic dependency resolves, filling the pipeline under the previous
        kernel's tail."""
        if not self._planned:
            raise RuntimeError("plan() must be called before run()")
        qlen = 1 if q_len_per_req is None else int(q_len_per_req)
        if qlen != getattr(self, "_qlen_planned", 1):
            bsz = self._plan_locals["last_page_len"].numel()
            if q.shape[0] != bsz * qlen:
                raise ValueError(f"q.shape[0] ({q.shape[0]}) does not match batch_size * q_len_per_req ({bsz} * {qlen} = {bsz * qlen}). "
                                 "For batch decode, q must have shape [batch_size * q_len_per_req, num_heads, head_dim].")
            self.plan(**{**self._plan_locals, "qo_indptr": None if qlen == 1 else torch.arange(0, (bsz + 1) * qlen, qlen, dtype=torch.int32)})
            self._qlen_planned = qlen
        if getattr(self, "_alibi_delegate", None) is not None:
            return self._alibi_delegate.run(q, paged_kv_cache, q_scale=q_scale, k_scale=k_scale, v_scale=v_scale, out=out, lse=lse,
                                            return_lse=return_lse, enable_pdl=enable_pdl, window_left=window_left, sinks=sinks)
        k_cache, v_cache = unpack_paged_kv_cache(paged_kv_cache, self._kv_layout)
        if getattr(self, "_rope", None) is not None:     # ROPE_LLAMA: rotate q and the batch's key pages, then the plain kernel
            from .attention.rope_on_the_fly import query_positions, rotate_rows, rotated_paged_keys

            qo = self._qo_indptr_host if self._qo_indptr_host is not None else torch.arange(self._batch_size + 1, dtype=torch.int32)
            q = rotate_rows(q, query_positions(qo, self._kv_lens_host), *self._rope)
            k_cache = rotated_paged_keys(k_cache, self._kv_indices, self._kv_indptr_host, self._kv_layout, *self._rope)
        if k_cache.dtype == torch.uint8 or v_cache.dtype == torch.uint8:
            # NVFP4 KV cache (reference decode.py:1293): block scales come in ``kv_cache_sf`` in the cache layout, the global
            # scales in k_scale / v_scale.  Composed path: the cache is widened to the query dtype, then the tcgen05 kernel runs.
            if kv_cache_sf is None:
                raise ValueError("kv_cache_sf must be provided for NVFP4 KV cache.")
            from .quantization.fp4 import nvfp4_dequantize_paged_kv_cache

            ksf, vsf = unpack_paged_kv_cache(kv_cache_sf, self._kv_layout)
            k_cache = nvfp4_dequantize_paged_kv_cache(k_cache, ksf, q.dtype)
            v_cache = nvfp4_dequantize_paged_kv_cache(v_cache, vsf, q.dtype)
        sm_scale = self._sm_scale
        if q_scale is not None:
            sm_scale *= q_scale
        if k_scale is not None:
            sm_scale *= k_scale
        window_left = self._window_left if window_left is None else window_left
        hq = self._num_qo_heads
        want_lse = return_lse or sinks is not None
        if out is None:
            out = torch.empty(q.shape[0], hq, self._head_dim, dtype=self._o_dtype, device=q.device)
        if want_lse and lse is None:
            lse = torch.empty(q.shape[0], hq, dtype=torch.float32, device=q.device)

        if not q.is_cuda:
            qo = self._qo_indptr_host
            if qo is None:
                qo = torch.arange(self._batch_size + 1, dtype=torch.int32)
            o_ref, lse_ref = reference.batch_paged_attention_ref(
                q, qo, k_cache, v_cache, self._kv_indptr_host, self._kv_indices.cpu(), self._kv_last_host,
                self._kv_layout, True, sm_scale, self._logits_soft_cap, window_left,
            )
            out.copy_(o_ref)
            if want_lse:
                lse.copy_(lse_ref)
        else:
            if self._run_sm100(q, k_cache, v_cache, out, lse if want_lse else None, sm_scale, window_left,
                               2 if (kv_prefetch and (enable_pdl is None or enable_pdl)) else enable_pdl, sinks=sinks):
                sinks = None  # folded into the softmax denominator by the kernel
        if sinks is not None:
            # the sink only adds exp(sink) to the softmax denominator: fold it in from (o, lse)
            from .attention._core import apply_attention_sink

            o2, l2 = apply_attention_sink(out, lse, sinks)
            out.copy_(o2)
            lse.copy_(l2)
        if v_scale is not None:
            out.copy_((out.float() * v_scale).to(out.dtype))
        return (out, lse) if return_lse else out

    def forward(self, q, paged_kv_cache, pos_encoding_mode="NONE", q_scale=None, k_scale=None, v_scale=None, window_left=-1, logits_soft_cap=None,
                sm_scale=None, rope_scale=None, rope_theta=None):
        """Deprecated (use :meth:`run`): the attention parameters given here replace the planned ones, defaults included."""
        legacy_forward_replan(self, pos_encoding_mode=pos_encoding_mode, window_left=window_left, logits_soft_cap=logits_soft_cap, sm_scale=sm_scale,
                              rope_scale=rope_scale, rope_theta=rope_theta)
        return self.run(q, paged_kv_cache, q_scale=q_scale, k_scale=k_scale, v_scale=v_scale)

    def forward_return_lse(self, q, paged_kv_cache, pos_encoding_mode="NONE", q_scale=None, k_scale=None, v_scale=None, window_left=-1,
                           logits_soft_cap=None, sm_scale=None, rope_scale=None, rope_theta=None):
        """Deprecated (use :meth:`run_return_lse`)."""
        legacy_forward_replan(self, pos_encoding_mode=pos_encoding_mode, window_left=window_left, logits_soft_cap=logits_soft_cap, sm_scale=sm_scale,
                              rope_scale=rope_scale, rope_theta=rope_theta)
        return self.run(q, paged_kv_cache, q_scale=q_scale, k_scale=k_scale, v_scale=v_scale, return_lse=True)

    def end_forward(self) -> None:
        pass

    def _run_generic(self, q, k_cache, v_cache, out, lse, sm_scale, window_left, enable_pdl):
        """Catch-all CUDA-core kernel: head dims other than 128, fp8 KV caches, very wide GQA groups."""
        from .attention import generic as _g

        if not _g.supported(q, k_cache, self._head_dim, self._head_dim):
            raise NotImplementedError(f"decode: unsupported configuration (q {q.dtype}, kv {k_cache.dtype}, head_dim {self._head_dim})")
        sp, sn, sh, page_size, hkv, d = paged_kv_strides(k_cache, self._kv_layout)
        vsp, vsn, vsh = paged_kv_strides(v_cache, self._kv_layout)[:3]
        dev = q.device
        if getattr(self, "_gen_key", None) != id(self._kv_indptr_host):
            qo = self._qo_indptr_host if self._qo_indptr_host is not None else torch.arange(self._batch_size + 1, dtype=torch.int32)
            self._gen_qo = qo.to(dev)
            self._gen_kv = self._kv_indptr_host.to(dev)
            self._gen_last = self._kv_last_host.to(dev)
            self._gen_key = id(self._kv_indptr_host)
        _g.run(q, k_cache, v_cache, out, lse, self._gen_qo, self._gen_kv, self._kv_indices, self._gen_last, page_size,
               (sp, sn, sh), (vsp, vsn, vsh), self._num_kv_heads, True, window_left, sm_scale, self._logits_soft_cap,
               enable_pdl=enable_pdl is None or enable_pdl)

    def _run_sm100(self, q, k_cache, v_cache, out, lse, sm_scale, window_left, enable_pdl, sinks=None):
        """Returns True when ``sinks`` were folded in by the kernel."""
        kv_ok = k_cache.dtype == q.dtype or k_cache.dtype in (torch.float8_e4m3fn, torch.float8_e5m2)
        if (self._head_dim != 128 or q.dtype not in (torch.float16, torch.bfloat16) or not kv_ok or v_cache.dtype != k_cache.dtype
                or self._max_q_rows > _MAX_Q_ROWS):
            self._run_generic(q, k_cache, v_cache, out, lse, sm_scale, window_left, enable_pdl)
            return False
        sp, sn, sh, page_size, hkv, d = paged_kv_strides(k_cache, self._kv_layout)
        if paged_kv_strides(v_cache, self._kv_layout)[:3] != (sp, sn, sh):
            raise ValueError("k_cache and v_cache must share strides")
        if page_size != self._page_size or hkv != self._num_kv_heads:
            raise ValueError("paged_kv_cache shape does not match plan()")
        if q.stride(-1) != 1 or out.stride(-1) != 1:
            raise ValueError("q/out last dim must be contiguous")
        causal = 1 if self._qo_indptr_host is not None else 0
        mod = jit.load("decode_sm100")
        mod.call(
            "decode_paged_run",
            q, k_cache, v_cache, out, lse, self._kv_indices, self._seg_info, self._cta_seg_indptr, self._merge_items,
            self._num_merge, self._partial_o, self._partial_lse, self._merge_counters, self._num_ctas, self._max_q_rows,
            self._num_qo_heads, self._num_kv_heads, self._head_dim, page_size, k_cache.shape[0], sp, sn, sh,
            1 if self._kv_layout == "HND" else 0, q.stride(0), q.stride(1), out.stride(0), out.stride(1),
            float(sm_scale), float(self._logits_soft_cap), int(window_left), causal,
            sinks.float().contiguous() if sinks is not None else None, dtype_code(q.dtype),
            dtype_code(k_cache.dtype), 2 if enable_pdl == 2 else (1 if (enable_pdl is None or enable_pdl) else 0), stream_ptr(q),
        )
        return sinks is not None


class CUDAGraphBatchDecodeWithPagedKVCacheWrapper(BatchDecodeWithPagedKVCacheWrapper):
    """CUDA-graph flavoured wrapper (fixed batch size, user-provided index buffers)."""

    def __init__(self, workspace_buffer, indptr_buffer, indices_buffer, last_page_len_buffer, kv_layout="NHD",
                 use_tensor_cores=True):
        super().__init__(
            workspace_buffer, kv_layout, use_cuda_graph=True, use_tensor_cores=use_tensor_cores,
            paged_kv_indptr_buffer=indptr_buffer, paged_kv_indices_buffer=indices_buffer,
            paged_kv_last_page_len_buffer=last_page_len_buffer,
        )


def fast_decode_plan(wrapper: BatchDecodeWithPagedKVCacheWrapper, *args, **kwargs) -> None:
    """Reference parity: flashinfer/decode.py:2897.  Our plan() is already a single C++ call plus
    one H2D copy, so the fast path is the same code."""
    wrapper.plan(*args, **kwargs)


# ------------------------------------------------------------------------------------------------
# Function-style decode APIs (block-table interface used by vLLM / TRT-LLM style engines).
# Parity: reference flashinfer/decode.py:2295-2724 (trtllm_batch_decode_with_kv_cache), :2727 (xqa_...),
# flashinfer/cudnn/decode.py:258.  There is one B200 kernel here, so they all route to it.
# ------------------------------------------------------------------------------------------------
def _block_tables_to_indices(block_tables: torch.Tensor, seq_lens: torch.Tensor, page_size: int):
    seq = seq_lens.to("cpu", torch.int64).reshape(-1)
    npages = (seq + page_size - 1) // page_size
    indptr = torch.zeros(seq.numel() + 1, dtype=torch.int32)
    indptr[1:] = npages.cumsum(0)
    bt = block_tables.to("cpu")
    if seq.numel():
        indices = torch.cat([bt[i, : int(npages[i])] for i in range(seq.numel())]).int()
    else:
        indices = torch.empty(0, dtype=torch.int32)
    last = torch.where(seq > 0, (seq - 1) % page_size + 1, torch.zeros_like(seq)).int()
    return indptr, indices, last


def trtllm_batch_decode_with_kv_cache(query: torch.Tensor, kv_cache, workspace_buffer: torch.Tensor,
                                      block_tables: torch.Tensor, seq_lens: torch.Tensor, max_seq_len: int,
                                      bmm1_scale: float = 1.0, bmm2_scale: float = 1.0, window_left: int = -1,
                                      out: Optional[torch.Tensor] = None, out_dtype=None, o_sf_scale=None,
                                      o_sf_vec_size=None, sinks=None, kv_layout: str = "HND", enable_pdl=None,
                                      backend: str = "auto", q_len_per_req: Optional[int] = 1, o_scale=None,
                                      mask=None, max_q_len=None, cum_seq_lens_q=None, skip_softmax_threshold_scale_factor=None,
                                      kv_cache_sf=None, uses_shared_paged_kv_idx: bool = True,
                                      lse=None, return_lse: bool = False):
    """``query [B * q_len_per_req, Hq, D]``; ``kv_cache`` a ``(k, v)`` tuple or ``[pages, 2, ...]`` tensor in
    ``kv_layout``; ``block_tables [B, max_pages]``; ``bmm1_scale`` is the softmax scale (q/k scales folded in)."""
    from .utils import reject_unsupported

    # NVFP4 output (o_sf_*), tree masks of speculative decoding and ragged query lengths are not implemented; the skip-softmax
    # threshold is a pure speed hint of the reference kernel (exact softmax here) and o_scale cancels in the result
    reject_unsupported("trtllm_batch_decode_with_kv_cache", o_sf_scale=o_sf_scale, o_sf_vec_size=o_sf_vec_size, mask=mask,
                       cum_seq_lens_q=cum_seq_lens_q, uses_shared_paged_kv_idx=(uses_shared_paged_kv_idx, True))
    k_cache, v_cache = unpack_paged_kv_cache(kv_cache, kv_layout)
    _, _, _, page_size, hkv, d = paged_kv_strides(k_cache, kv_layout)
    if k_cache.dtype == torch.uint8:
        d *= 2  # NVFP4 cache: two e2m1 values per byte
    indptr, indices, last = _block_tables_to_indices(block_tables, seq_lens, page_size)
    b = seq_lens.numel()
    w = BatchDecodeWithPagedKVCacheWrapper(workspace_buffer, kv_layout)
    ql = q_len_per_req or 1
    qo = torch.arange(0, (b + 1) * ql, ql, dtype=torch.int32) if ql > 1 else None
    w.plan(indptr, indices, last, query.shape[1], hkv, d, page_size, window_left=window_left, q_data_type=query.dtype,
           sm_scale=float(bmm1_scale), qo_indptr=qo)
    res = w.run(query, (k_cache, v_cache), out=out if (out is not None and out.dtype == query.dtype) else None,
                return_lse=return_lse, sinks=sinks, v_scale=float(bmm2_scale) if float(bmm2_scale) != 1.0 else None,
                kv_cache_sf=kv_cache_sf)
    if return_lse:
        o, l = res
        if lse is not None:
            lse.copy_(l.view_as(lse))
            l = lse
        if out is not None and o.data_ptr() != out.data_ptr():
            out.copy_(o)
            o = out
        elif out_dtype is not None and o.dtype != out_dtype:
            o = o.to(out_dtype)
        return o, l
    if out is not None and res.data_ptr() != out.data_ptr():
        out.copy_(res)
        return out
    if out is None and out_dtype is not None and res.dtype != out_dtype:
        return res.to(out_dtype)
    return res


def xqa_batch_decode_with_kv_cache(query: torch.Tensor, kv_cache, workspace_buffer: torch.Tensor, block_tables: torch.Tensor,
                                   seq_lens: torch.Tensor, max_seq_len: int, bmm1_scale: float = 1.0, bmm2_scale: float = 1.0,
                                   window_left: int = -1, out: Optional[torch.Tensor] = None, sinks=None, kv_layout: str = "NHD",
                                   enable_pdl=None, q_len_per_req: Optional[int] = 1, o_scale: Optional[float] = 1.0, mask=None,
                                   kv_cache_sf=None):
    """XQA entry point of the reference (flashinfer/decode.py:2727; note its ``kv_layout`` default is NHD): same kernel here.
    ``o_scale`` cancels in the result (the reference multiplies the V scale by it and the output by its reciprocal; it only
    positions the fp8 output range of that kernel), so it is accepted and not applied."""
    return trtllm_batch_decode_with_kv_cache(query, kv_cache, workspace_buffer, block_tables, seq_lens, max_seq_len, bmm1_scale,
                                             bmm2_scale, window_left, out=out, sinks=sinks, kv_layout=kv_layout,
                                             enable_pdl=enable_pdl, q_len_per_req=q_len_per_req, mask=mask, kv_cache_sf=kv_cache_sf)


def cudnn_batch_decode_with_kv_cache(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, scale: float,
                                     workspace_buffer: torch.Tensor, *, max_sequence_kv: int,
                                     actual_seq_lens_kv: Optional[torch.Tensor] = None,
                                     block_tables: Optional[torch.Tensor] = None, is_cuda_graph_compatible: bool = False,
                                     batch_offsets_q=None, batch_offsets_o=None, batch_offsets_k=None,
                                     batch_offsets_v=None, out: Optional[torch.Tensor] = None):
    """cuDNN-style decode signature (reference flashinfer/cudnn/decode.py:258): HND paged caches + block tables."""
    return trtllm_batch_decode_with_kv_cache(q, (k_cache, v_cache), workspace_buffer, block_tables,
                                             actual_seq_lens_kv.reshape(-1), max_sequence_kv, bmm1_scale=scale, out=out,
                                             kv_layout="HND")


class BatchDecodeMlaWithPagedKVCacheWrapper:
    """Legacy MLA decode wrapper (reference flashinfer/decode.py:1677-2054): q_nope/q_pe + compressed-kv / k-pe caches."""

    def __init__(self, float_workspace_buffer: torch.Tensor, use_cuda_graph: bool = False, use_tensor_cores: bool = True,
                 paged_kv_indptr_buffer=None, paged_kv_indices_buffer=None, paged_kv_last_page_len_buffer=None) -> None:
        from .mla import BatchMLAPagedAttentionWrapper

        self._w = BatchMLAPagedAttentionWrapper(float_workspace_buffer, use_cuda_graph)
        self._use_cuda_graph = bool(use_cuda_graph)

    @property
    def is_cuda_graph_enabled(self) -> bool:
        return self._use_cuda_graph

    @property
    def use_tensor_cores(self) -> bool:
        return True  # the MLA kernel is tcgen05

    def reset_workspace_buffer(self, float_workspace_buffer: torch.Tensor, int_workspace_buffer: Optional[torch.Tensor] = None) -> None:
        self._w._float_workspace_buffer = float_workspace_buffer

    def plan(self, indptr, indices, last_page_len, num_qo_heads, head_dim_compressed_kv, page_size, sm_scale,
             window_left: int = -1, logits_soft_cap=None, data_type="float16", q_data_type=None, rope_scale=None,
             rope_theta=None) -> None:
        if window_left is not None and window_left >= 0:
            raise NotImplementedError("BatchDecodeMlaWithPagedKVCacheWrapper: sliding windows are not implemented by the MLA kernel")
        if logits_soft_cap:
            raise NotImplementedError("BatchDecodeMlaWithPagedKVCacheWrapper: logits_soft_cap is not implemented by the MLA kernel")
        n_pages = (indptr[1:] - indptr[:-1]).to("cpu", torch.int64)
        kv_len = torch.clamp(n_pages - 1, min=0) * page_size + last_page_len.to("cpu", torch.int64)
        b = kv_len.numel()
        dt = _canon_dtype(q_data_type or data_type)
        self._plan_args, self._sm_scale = (indptr, indices, kv_len.int(), num_qo_heads, head_dim_compressed_kv, page_size, dt, b), float(sm_scale)
        self._sm_planned = self._sm_scale
        self._w.plan(torch.arange(b + 1, dtype=torch.int32), indptr, indices, kv_len.int(), num_qo_heads,
                     head_dim_compressed_kv, 64, page_size, False, sm_scale, dt, dt)

    begin_forward = plan

    def run(self, q_nope, q_pe, paged_ckv_cache, paged_kpe_cache, q_scale=None, k_scale=None, v_scale=None, out=None,
            lse=None, return_lse: bool = False, enable_pdl: bool = False):
        # q / k de-quantisation scales fold into the softmax scale (a re-plan of the host-side work list when they change the
        # planned value); the v scale multiplies the output
        sm = self._sm_scale * (float(q_scale) if q_scale is not None else 1.0) * (float(k_scale) if k_scale is not None else 1.0)
        if sm != getattr(self, "_sm_planned", self._sm_scale):
            indptr, indices, kv_len, h, dckv, page_size, dt, b = self._plan_args
            self._w.plan(torch.arange(b + 1, dtype=torch.int32), indptr, indices, kv_len, h, dckv, 64, page_size, False, sm, dt, dt)
            self._sm_planned = sm
        return self._w.run(q_nope, q_pe, paged_ckv_cache, paged_kpe_cache, out=out, lse=lse, return_lse=return_lse,
                           o_scale=float(v_scale) if (v_scale is not None and float(v_scale) != 1.0) else None)

    forward = run


from . import jit as _jit_acc  # noqa: E402

get_batch_decode_module = _jit_acc.module_accessor("decode_sm100")
get_batch_decode_jit_module = _jit_acc.module_accessor("decode_sm100")
get_single_decode_module = _jit_acc.module_accessor("decode_sm100")
get_batch_decode_mla_module = _jit_acc.module_accessor("mla_sm100")
get_trtllm_gen_decode_module = _jit_acc.module_accessor("decode_sm100")
get_trtllm_gen_fmha_module = _jit_acc.module_accessor("decode_sm100")


def single_decode_with_kv_cache_with_jit_module(jit_module, q, k, v, *args, kv_layout: str = "NHD", window_left: int = -1,
                                                return_lse: bool = False, **kwargs):
    """Reference decode.py: run single-request decode through an explicitly supplied JIT module (custom attention variants).
    Variants here live in the generic attention module and are selected by arguments, so ``jit_module`` only has to be a loaded
    native module; the call is :func:`single_decode_with_kv_cache`."""
    return single_decode_with_kv_cache(q, k, v, kv_layout=kv_layout, window_left=window_left, return_lse=return_lse, **kwargs)
