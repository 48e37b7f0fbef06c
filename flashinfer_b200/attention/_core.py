"""``BatchAttention``: one plan()/run() for arbitrary mixes of prefill / append / decode requests.
Parity: reference flashinfer/attention/_core.py:44-216 (persistent two-runner kernel) and the attention-sink wrapper (:218).

Here the planner splits the batch by packed query rows: requests with ``q_len * group <= 32`` go to the swap-AB decode
kernel, the rest to the FMHA prefill kernel; both are persistent and are co-scheduled on disjoint SM budgets (see pod.py).
Attention sinks are folded in afterwards from (o, lse): ``o' = o * sigmoid(lse_e - sink)`` (the sink only adds
``exp(sink)`` to the softmax denominator).
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import torch

from ..decode import BatchDecodeWithPagedKVCacheWrapper
from ..pod import _SideStream, _split_sms
from ..prefill import BatchPrefillWithPagedKVCacheWrapper
from ..utils import device_sm_count

LOG2E = 1.4426950408889634


def apply_attention_sink(out: torch.Tensor, lse: torch.Tensor, sinks: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fold per-head sink logits into an attention state.  ``lse`` is base-2; ``sinks [H]`` are natural-log logits."""
    s2 = sinks.float()[None, :] * LOG2E
    new_lse = torch.logaddexp2(lse.float(), s2) if hasattr(torch, "logaddexp2") else torch.log2(torch.exp2(lse) + torch.exp2(s2))
    scale = torch.exp2(lse.float() - new_lse)
    return (out.float() * scale[..., None]).to(out.dtype), new_lse


class BatchAttention:
    def __init__(self, kv_layout: str = "NHD", device: str = "cuda") -> None:
        self._kv_layout = kv_layout
        self.device = torch.device(device)
        self._ws = torch.empty(256 << 20, dtype=torch.uint8, device=self.device)
        half = self._ws.numel() // 2
        self._prefill = BatchPrefillWithPagedKVCacheWrapper(self._ws[:half], kv_layout)
        self._decode = BatchDecodeWithPagedKVCacheWrapper(self._ws[half:], kv_layout)
        self._side = _SideStream(self.device)

    def plan(self, qo_indptr: torch.Tensor, kv_indptr: torch.Tensor, kv_indices: torch.Tensor, kv_len_arr: torch.Tensor,
             num_qo_heads: int, num_kv_heads: int, head_dim_qk: int, head_dim_vo: int, page_size: int,
             causal: bool = False, sm_scale: float = None, logits_soft_cap: Optional[float] = None,
             q_data_type: torch.dtype = torch.bfloat16, kv_data_type: torch.dtype = torch.bfloat16,
             use_profiler: bool = False) -> None:
        qo = qo_indptr.to("cpu", torch.int64)
        kvp = kv_indptr.to("cpu", torch.int64)
        kvl = kv_len_arr.to("cpu", torch.int64)
        idx = kv_indices.to("cpu", torch.int32)
        # plan()-time state read by the fi_trace template of run()
        self._qo_indptr_host, self._kv_indptr_host, self._kv_len_host, self._kv_indices_host = qo.int(), kvp.int(), kvl.int(), idx
        self._causal = bool(causal)
        self._sm_scale = float(sm_scale) if sm_scale is not None else 1.0 / (head_dim_qk ** 0.5)
        group = num_qo_heads // num_kv_heads
        q_lens = qo[1:] - qo[:-1]
        is_dec = (q_lens * group <= 32) & (q_lens > 0) & bool(causal or (q_lens == 1).all())
        self._sel_d = torch.nonzero(is_dec).flatten()
        self._sel_p = torch.nonzero(~is_dec & (q_lens > 0)).flatten()
        self._hq, self._dvo = num_qo_heads, head_dim_vo
        self._soft_cap = float(logits_soft_cap or 0.0)

        def sub(sel):
            qi, kpi, ind, last, rows = [0], [0], [], [], []
            for b in sel.tolist():
                n_pages = int(kvp[b + 1] - kvp[b])
                used = min(n_pages, (int(kvl[b]) + page_size - 1) // page_size)
                ind.append(idx[int(kvp[b]) : int(kvp[b]) + used])
                kpi.append(kpi[-1] + used)
                last.append(int(kvl[b]) - (used - 1) * page_size if used > 0 else 0)
                qi.append(qi[-1] + int(q_lens[b]))
                rows.append(torch.arange(int(qo[b]), int(qo[b + 1])))
            cat = lambda xs, dt: torch.cat(xs).to(dt) if xs else torch.empty(0, dtype=dt)  # noqa: E731
            return (torch.tensor(qi, dtype=torch.int32), torch.tensor(kpi, dtype=torch.int32), cat(ind, torch.int32),
                    torch.tensor(last, dtype=torch.int32), cat(rows, torch.int64))

        total = device_sm_count(self.device if self.device.type == "cuda" else None)
        flops = float(sum(4.0 * int(q_lens[b]) * int(kvl[b]) for b in self._sel_p.tolist())) * num_qo_heads * head_dim_qk
        dbytes = float(sum(int(kvl[b]) for b in self._sel_d.tolist())) * num_kv_heads * head_dim_qk * 4
        sp, sd = _split_sms(total, flops, dbytes)
        self._prefill._cta_budget, self._decode._cta_budget = sp or None, sd or None
        self._rows_p = self._rows_d = None
        if self._sel_p.numel():
            qi, kpi, ind, last, rows = sub(self._sel_p)
            self._prefill.plan(qi, kpi, ind, last, num_qo_heads, num_kv_heads, head_dim_qk, page_size, causal=causal,
                               sm_scale=sm_scale, logits_soft_cap=logits_soft_cap, q_data_type=q_data_type,
                               kv_data_type=kv_data_type)
            self._rows_p = rows.to(self.device)
        if self._sel_d.numel():
            qi, kpi, ind, last, rows = sub(self._sel_d)
            plain = bool((q_lens[self._sel_d] == 1).all())
            self._decode.plan(kpi, ind, last, num_qo_heads, num_kv_heads, head_dim_qk, page_size, sm_scale=sm_scale,
                              logits_soft_cap=logits_soft_cap, q_data_type=q_data_type, kv_data_type=kv_data_type,
                              qo_indptr=None if plain else qi)
            self._rows_d = rows.to(self.device)

    def run(self, q: torch.Tensor, kv_cache, out: Optional[torch.Tensor] = None, lse: Optional[torch.Tensor] = None,
            k_scale=None, v_scale=None, logits_soft_cap: float = 0.0, profiler_buffer=None, kv_cache_sf=None):
        from ..utils import reject_unsupported

        reject_unsupported("BatchAttention.run", kv_cache_sf=kv_cache_sf)
        if logits_soft_cap and float(logits_soft_cap) != float(getattr(self, "_soft_cap", 0.0) or 0.0):
            raise ValueError("BatchAttention.run: logits_soft_cap differs from the value given to plan() (it is compiled into the plan)")
        scales = {k_: v_ for k_, v_ in (("k_scale", k_scale), ("v_scale", v_scale)) if v_ is not None}
        if out is None:
            out = torch.empty(q.shape[0], self._hq, self._dvo, dtype=q.dtype, device=q.device)
        if lse is None:
            lse = torch.empty(q.shape[0], self._hq, dtype=torch.float32, device=q.device)
        both = self._rows_p is not None and self._rows_d is not None and q.is_cuda
        if both:
            self._side.fork()
        if self._rows_d is not None:
            ctx = torch.cuda.stream(self._side.stream) if both else _null()
            with ctx:
                o_d, l_d = self._decode.run(q[self._rows_d], kv_cache, return_lse=True, **scales)
                out[self._rows_d] = o_d
                lse[self._rows_d] = l_d
        if self._rows_p is not None:
            o_p, l_p = self._prefill.run(q[self._rows_p], kv_cache, return_lse=True, **scales)
            out[self._rows_p] = o_p
            lse[self._rows_p] = l_p
        if both:
            self._side.join()
        return out, lse


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class BatchAttentionWithAttentionSinkWrapper(BatchPrefillWithPagedKVCacheWrapper):
    """Paged prefill/decode with per-head attention sinks (gpt-oss style)."""

    def __init__(self, float_workspace_buffer: torch.Tensor, kv_layout: str = "NHD", use_cuda_graph: bool = False,
                 **kwargs) -> None:
        kwargs.pop("jit_args", None)
        kwargs.pop("jit_kwargs", None)
        kwargs.pop("backend", None)
        kwargs.pop("q_data_type", None)
        kwargs.pop("kv_data_type", None)
        kwargs.pop("head_dim_qk", None)
        kwargs.pop("head_dim_vo", None)
        kwargs.pop("window_left", None)
        super().__init__(float_workspace_buffer, kv_layout, use_cuda_graph)

    def run(self, q, paged_kv_cache, sinks: Optional[torch.Tensor] = None, sm_scale=None, *args, return_lse=False, **kw):
        if sm_scale is not None:
            self._sm_scale = float(sm_scale)          # the reference passes the softmax scale at run() time for this wrapper
        o, l = super().run(q, paged_kv_cache, return_lse=True, **kw)
        if sinks is not None:
            o, l = apply_attention_sink(o, l, sinks)
        return (o, l) if return_lse else o
