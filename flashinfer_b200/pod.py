"""POD-Attention: prefill and decode co-scheduled on disjoint SM sets.

Parity: reference flashinfer/pod.py:61-1205 (PODWithPagedKVCacheWrapper, BatchPODWithPagedKVCacheWrapper) and
include/flashinfer/attention/pod.cuh (one grid, CTAs pick PREFILL/DECODE by SM id at run time).

B200-first design: both attention kernels are persistent (one CTA per SM, grid size chosen by the planner), so SM
sharing is decided on the host: a cost model (prefill FLOPs at tensor peak vs. decode bytes at HBM peak) splits the
148 SMs into two budgets, each phase is planned for its budget and the two kernels are launched on forked streams —
they run concurrently on disjoint SMs, the memory-bound decode hiding under the compute-bound prefill.
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import torch

from . import jit
from .decode import BatchDecodeWithPagedKVCacheWrapper
from .prefill import BatchPrefillWithPagedKVCacheWrapper, BatchPrefillWithRaggedKVCacheWrapper
from .utils import check_pos_encoding_mode, device_sm_count

_PEAK_FLOPS = 1.1e15  # achieved FMHA rate (profiles/), not the datasheet number
_PEAK_BW = 6.0e12


def _split_sms(total: int, prefill_flops: float, decode_bytes: float) -> Tuple[int, int]:
    t_p, t_d = prefill_flops / _PEAK_FLOPS, decode_bytes / _PEAK_BW
    if t_p <= 0:
        return 0, total
    if t_d <= 0:
        return total, 0
    p = int(round(total * t_p / (t_p + t_d)))
    p = max(8, min(total - 8, p))
    return p, total - p


def _fusable(q_p: torch.Tensor, q_d: torch.Tensor, kv_d, hq: int, hkv: int, head_dim: int) -> bool:
    """Conditions of the single-kernel path (csrc/attention/pod_sm100.cu): 16-bit q of one dtype on both sides, 16-bit decode
    KV cache, head_dim 128, decode rows per KV head (GQA group, q_len 1) in {1, 4, 8}."""
    kd = kv_d[0] if isinstance(kv_d, (tuple, list)) else kv_d
    return (q_p.is_cuda and q_p.dtype in (torch.float16, torch.bfloat16) and q_d.dtype == q_p.dtype and kd.dtype == q_p.dtype
            and head_dim == 128 and hq % hkv == 0 and (hq // hkv) in (1, 4, 8) and q_d.shape[0] > 0 and q_p.shape[0] > 0)


def _run_fused(prefill_call, decode_call):
    """ONE launch for both phases: with the stage armed, the prefill launcher parks its launch inside ``pod_sm100`` and the
    decode launcher that follows fuses both into ``pod_kernel`` (prefill CTAs first, decode CTAs after them).  If the
    prefill side did not take the tcgen05 path (nothing parked) the decode simply runs after it."""
    mod = jit.load("pod_sm100")
    flag = torch.zeros(1, dtype=torch.int64)
    mod.call("pod_arm", 1)
    try:
        with jit.redirect({"prefill_sm100": "pod_sm100", "decode_sm100": "pod_sm100"}):
            res_p = prefill_call()
            mod.call("pod_query", flag)
            if int(flag[0]) == 0:
                mod.call("pod_arm", 0)
            res_d = decode_call()
    finally:
        mod.call("pod_arm", 0)
    return res_p, res_d


class _SideStream:
    def __init__(self, device):
        self.stream = torch.cuda.Stream(device=device) if device.type == "cuda" else None

    def fork(self):
        ev = torch.cuda.Event()
        ev.record()
        self.stream.wait_event(ev)

    def join(self):
        ev = torch.cuda.Event()
        ev.record(self.stream)
        torch.cuda.current_stream().wait_event(ev)


class PODWithPagedKVCacheWrapper:
    """One prefill request (contiguous k_p/v_p) fused with a batch of paged decodes."""

    def __init__(self, float_workspace_buffer: torch.Tensor, kv_layout: str = "NHD", use_cuda_graph: bool = False,
                 paged_kv_indptr_buffer=None, paged_kv_indices_buffer=None, paged_kv_last_page_len_buffer=None,
                 jit_args=None) -> None:
        self.device = float_workspace_buffer.device
        half = float_workspace_buffer.numel() // 2
        self._decode = BatchDecodeWithPagedKVCacheWrapper(float_workspace_buffer[half:], kv_layout)
        self._prefill = BatchPrefillWithRaggedKVCacheWrapper(float_workspace_buffer[:half], kv_layout)
        self._kv_layout = kv_layout
        self._side = _SideStream(self.device)
        self._fused = True  # single-kernel POD when the shapes allow it (set False to force the two-stream composition)
        self._use_cuda_graph = bool(use_cuda_graph)

    @property
    def is_cuda_graph_enabled(self) -> bool:
        return self._use_cuda_graph

    def reset_workspace_buffer(self, float_workspace_buffer: torch.Tensor, int_workspace_buffer: Optional[torch.Tensor] = None) -> None:
        half = float_workspace_buffer.numel() // 2
        self._prefill.reset_workspace_buffer(float_workspace_buffer[:half], int_workspace_buffer)
        self._decode.reset_workspace_buffer(float_workspace_buffer[half:], int_workspace_buffer)

    def plan(self, indptr, indices, last_page_len, num_qo_heads, num_kv_heads, head_dim, page_size,
             pos_encoding_mode="NONE", window_left=-1, q_data_type="float16", kv_data_type=None, data_type=None,
             sm_scale=None, rope_scale=None, rope_theta=None, non_blocking=True) -> None:
        check_pos_encoding_mode(pos_encoding_mode)
        # like in the reference, the decode side uses the positional encoding recorded here (run(pos_encoding_mode_d=) is not consulted)
        self._dargs = (indptr, indices, last_page_len, num_qo_heads, num_kv_heads, head_dim, page_size)
        self._dkw = dict(window_left=window_left, q_data_type=data_type or q_data_type, kv_data_type=kv_data_type,
                         sm_scale=sm_scale, pos_encoding_mode=pos_encoding_mode, rope_scale=rope_scale, rope_theta=rope_theta)
        self._plain = pos_encoding_mode == "NONE"
        self._hq, self._hkv, self._d, self._ps = num_qo_heads, num_kv_heads, head_dim, page_size
        n_pages = (indptr[1:] - indptr[:-1]).to("cpu")
        self._decode_tokens = int(n_pages.sum()) * page_size
        self._decode_planned_for = None

    begin_forward = plan

    def run(self, q_p, k_p, v_p, q_d, paged_kv_cache_d, custom_mask_p=None, packed_custom_mask_p=None, causal_p=False,
            kv_layout_p="NHD", pos_encoding_mode_p="NONE", sm_scale_p=None, window_left_p=-1, rope_scale_p=None,
            rope_theta_p=None, return_lse_p=False, custom_mask_d=None, packed_custom_mask_d=None, causal_d=False,
            kv_layout_d="NHD", pos_encoding_mode_d="NONE", sm_scale_d=None, window_left_d=-1, rope_scale_d=None,
            rope_theta_d=None, q_scale=None, k_scale=None, v_scale=None, return_lse_d=False,
            use_fp16_qk_reduction=False, enable_pdl=None):
        # custom_mask_d / packed_custom_mask_d / pos_encoding_mode_d are accepted and not consulted, as in the reference (one query
        # token per request sees its whole history; the decode positional encoding is the one given to plan()).  A custom prefill
        # mask or a positional encoding on either side runs the two phases as two launches (side stream) - the single-launch
        # fusion covers the plain kernels
        check_pos_encoding_mode(pos_encoding_mode_p)
        masked = custom_mask_p is not None or packed_custom_mask_p is not None
        plain = self._plain and pos_encoding_mode_p == "NONE" and not masked
        # kv_layout_d is not consulted: like in the reference, the decode cache has the layout the wrapper was constructed with
        # run()-time decode parameters override what plan() recorded (the decode side is planned lazily, below); causal_d is
        # immaterial for one query token per request
        dkw = dict(self._dkw)
        if sm_scale_d is not None:
            dkw["sm_scale"] = sm_scale_d
        if window_left_d != -1:
            dkw["window_left"] = window_left_d
        if dkw != self._dkw:
            self._dkw, self._decode_planned_for = dkw, None
        total = device_sm_count(self.device if self.device.type == "cuda" else None)
        qo_len, kv_len = q_p.shape[0], (k_p.shape[0] if kv_layout_p == "NHD" else k_p.shape[1])
        flops = 4.0 * qo_len * kv_len * self._hq * self._d * (0.5 if causal_p else 1.0)
        dbytes = 2.0 * self._decode_tokens * self._hkv * self._d * 2
        sp, sd = _split_sms(total, flops, dbytes)
        self._prefill._cta_budget, self._decode._cta_budget = sp or None, sd or None
        if self._decode_planned_for != sd:
            self._decode.plan(*self._dargs, **self._dkw)
            self._decode_planned_for = sd
        self._prefill._kv_layout = kv_layout_p
        self._prefill.plan(torch.tensor([0, qo_len], dtype=torch.int32), torch.tensor([0, kv_len], dtype=torch.int32),
                           self._hq, self._hkv, self._d, causal=causal_p and not masked, sm_scale=sm_scale_p,
                           window_left=window_left_p, q_data_type=q_p.dtype, custom_mask=custom_mask_p,
                           packed_custom_mask=packed_custom_mask_p, pos_encoding_mode=pos_encoding_mode_p,
                           rope_scale=rope_scale_p, rope_theta=rope_theta_p)
        if plain and _fusable(q_p, q_d, paged_kv_cache_d, self._hq, self._hkv, self._d) and k_p.dtype == q_p.dtype \
                and q_scale is None and k_scale is None and v_scale is None and self._fused:
            res_p, res_d = _run_fused(lambda: self._prefill.run(q_p, k_p, v_p, return_lse=return_lse_p),
                                      lambda: self._decode.run(q_d, paged_kv_cache_d, return_lse=return_lse_d))
        elif q_p.is_cuda:
            self._side.fork()
            with torch.cuda.stream(self._side.stream):
                res_d = self._decode.run(q_d, paged_kv_cache_d, q_scale=q_scale, k_scale=k_scale, v_scale=v_scale,
                                         return_lse=return_lse_d)
            res_p = self._prefill.run(q_p, k_p, v_p, return_lse=return_lse_p)
            self._side.join()
        else:
            res_d = self._decode.run(q_d, paged_kv_cache_d, return_lse=return_lse_d)
            res_p = self._prefill.run(q_p, k_p, v_p, return_lse=return_lse_p)
        return res_p, res_d

    forward = run

    def end_forward(self) -> None:
        pass


class BatchPODWithPagedKVCacheWrapper:
    """Batched prefill (paged) fused with batched decode (paged)."""

    def __init__(self, float_workspace_buffer: torch.Tensor, kv_layout: str = "NHD", use_cuda_graph: bool = False,
                 **kwargs) -> None:
        self.device = float_workspace_buffer.device
        half = float_workspace_buffer.numel() // 2
        self._prefill = BatchPrefillWithPagedKVCacheWrapper(float_workspace_buffer[:half], kv_layout)
        self._decode = BatchDecodeWithPagedKVCacheWrapper(float_workspace_buffer[half:], kv_layout)
        self._side = _SideStream(self.device)
        self._fused = True
        self._use_cuda_graph = bool(use_cuda_graph)

    @property
    def is_cuda_graph_enabled(self) -> bool:
        return self._use_cuda_graph

    def plan(self, qo_indptr_p, kv_indptr_p, kv_indices_p, last_page_len_p, qo_indptr_d, kv_indptr_d, kv_indices_d,
             last_page_len_d, num_qo_heads, num_kv_heads, head_dim, page_size, pos_encoding_mode="NONE",
             window_left=-1, q_data_type="float16", kv_data_type=None, data_type=None, sm_scale=None, rope_scale=None,
             rope_theta=None, non_blocking=True, causal_p: Optional[bool] = None) -> None:
        """``causal_p``: the reference chooses the prefill mask at run() time (``run(causal_p=...)``, default non-causal).  The tcgen05
        prefill plan bakes the mask in, so it can also be given here; run() re-plans the prefill side when it asks for the other
        mask.  Left at None in both places, the reference's default (non-causal) applies."""
        check_pos_encoding_mode(pos_encoding_mode)
        self._plain = pos_encoding_mode == "NONE"
        self._causal_arg = causal_p
        causal_p = True if causal_p is None else bool(causal_p)       # plan for the common case; run() corrects it if needed
        self._causal_planned = causal_p
        total = device_sm_count(self.device if self.device.type == "cuda" else None)
        qo_p = qo_indptr_p.to("cpu", torch.int64)
        q_lens = qo_p[1:] - qo_p[:-1]
        np_p = (kv_indptr_p[1:] - kv_indptr_p[:-1]).to("cpu", torch.int64)
        kv_lens_p = np_p * page_size
        flops = float((4.0 * q_lens * kv_lens_p).sum()) * num_qo_heads * head_dim * (0.5 if causal_p else 1.0)
        dbytes = 2.0 * float((kv_indptr_d[1:] - kv_indptr_d[:-1]).sum()) * page_size * num_kv_heads * head_dim * 2
        sp, sd = _split_sms(total, flops, dbytes)
        self._prefill._cta_budget, self._decode._cta_budget = sp or None, sd or None
        dt = data_type or q_data_type
        self._pplan = ((qo_indptr_p, kv_indptr_p, kv_indices_p, last_page_len_p, num_qo_heads, num_kv_heads, head_dim, page_size),
                       dict(sm_scale=sm_scale, window_left=window_left, q_data_type=dt, kv_data_type=kv_data_type,
                            pos_encoding_mode=pos_encoding_mode, rope_scale=rope_scale, rope_theta=rope_theta))
        self._mask_planned = None
        self._prefill.plan(*self._pplan[0], causal=causal_p, **self._pplan[1])
        qo_d = qo_indptr_d.to("cpu", torch.int64)
        plain_decode = bool(((qo_d[1:] - qo_d[:-1]) == 1).all())
        self._decode.plan(kv_indptr_d, kv_indices_d, last_page_len_d, num_qo_heads, num_kv_heads, head_dim, page_size,
                          window_left=window_left, q_data_type=dt, kv_data_type=kv_data_type, sm_scale=sm_scale,
                          qo_indptr=None if plain_decode else qo_indptr_d, pos_encoding_mode=pos_encoding_mode,
                          rope_scale=rope_scale, rope_theta=rope_theta)
        self._sm_split = (sp, sd)
        self._hq, self._hkv, self._d, self._plain_decode = num_qo_heads, num_kv_heads, head_dim, plain_decode

    begin_forward = plan

    def run(self, q_p, paged_kv_cache_p, q_d, paged_kv_cache_d, custom_mask_p=None, packed_custom_mask_p=None,
            causal_p: Optional[bool] = None, q_scale=None, k_scale=None, v_scale=None, return_lse: bool = False,
            use_fp16_qk_reduction: bool = False, enable_pdl=None):
        want = bool(causal_p) if causal_p is not None else (bool(self._causal_arg) if self._causal_arg is not None else False)
        mask = custom_mask_p if custom_mask_p is not None else packed_custom_mask_p
        if mask is not None:                              # the prefill plan owns the mask (as in the prefill wrappers): re-plan with it
            if self._mask_planned is not mask:
                self._prefill.plan(*self._pplan[0], causal=False, custom_mask=custom_mask_p,
                                   packed_custom_mask=None if custom_mask_p is not None else packed_custom_mask_p, **self._pplan[1])
                self._mask_planned, self._causal_planned = mask, None
        elif want != self._causal_planned:
            self._prefill.plan(*self._pplan[0], causal=want, **self._pplan[1])
            self._causal_planned, self._mask_planned = want, None
        kp = paged_kv_cache_p[0] if isinstance(paged_kv_cache_p, (tuple, list)) else paged_kv_cache_p
        if self._fused and self._plain and mask is None and self._plain_decode and _fusable(q_p, q_d, paged_kv_cache_d, self._hq, self._hkv, self._d) \
                and kp.dtype == q_p.dtype and q_scale is None and k_scale is None and v_scale is None:
            res_p, res_d = _run_fused(lambda: self._prefill.run(q_p, paged_kv_cache_p, return_lse=return_lse),
                                      lambda: self._decode.run(q_d, paged_kv_cache_d, return_lse=return_lse))
        elif q_p.is_cuda:
            self._side.fork()
            with torch.cuda.stream(self._side.stream):
                res_d = self._decode.run(q_d, paged_kv_cache_d, q_scale=q_scale, k_scale=k_scale, v_scale=v_scale,
                                         return_lse=return_lse)
            res_p = self._prefill.run(q_p, paged_kv_cache_p, return_lse=return_lse)
            self._side.join()
        else:
            res_d = self._decode.run(q_d, paged_kv_cache_d, return_lse=return_lse)
            res_p = self._prefill.run(q_p, paged_kv_cache_p, return_lse=return_lse)
        return res_p, res_d

    forward = run

    def end_forward(self) -> None:
        pass


def get_pod_module(*args, **kwargs):
    """The native module behind this file's ops (reference pod.py get_pod_module: the JIT module accessor)."""
    from . import jit

    return jit.load("pod_sm100")


def get_batch_pod_module(*args, **kwargs):
    """The native module behind this file's ops (reference pod.py get_batch_pod_module: the JIT module accessor)."""
    from . import jit

    return jit.load("pod_sm100")
