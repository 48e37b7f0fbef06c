"""MLA paged attention (DeepSeek-V2/V3).  Parity: reference flashinfer/mla/_core.py:219-940.

CUDA tensors run the tcgen05 kernel ``csrc/attention/mla_sm100.cu`` (all heads of a query token on MMA-M, the
256-wide halves of d_v on two CTAs, split-KV + merge); CPU tensors run the fp32 oracle.
"""
from __future__ import annotations

import dataclasses

import math
from typing import List, Optional, Tuple, Union

import torch

from .. import cascade, jit
from ..utils import device_sm_count, dtype_code, stream_ptr

_TILE = 32
_WORK_INTS = 8
LOG2E = 1.4426950408889634


def mla_attention_ref(q_nope, q_pe, ckv, kpe, sm_scale, causal_offset: Optional[int] = None):
    """fp32 oracle for one request: q_nope [q, H, 512], q_pe [q, H, 64], ckv [kv, 512], kpe [kv, 64].
    Query i sees kv positions <= kv_len - q_len + i when ``causal_offset`` is None -> all."""
    ql, kv_len = q_nope.shape[0], ckv.shape[0]
    logits = (torch.einsum("qhd,kd->hqk", q_nope.float(), ckv.float()) +
              torch.einsum("qhd,kd->hqk", q_pe.float(), kpe.float())) * sm_scale
    if causal_offset is not None:
        qpos = torch.arange(ql, device=q_nope.device)[:, None] + causal_offset
        mask = torch.arange(kv_len, device=q_nope.device)[None, :] <= qpos
        logits = logits.masked_fill(~mask[None], float("-inf"))
    lse = torch.logsumexp(logits, -1)
    p = torch.exp(logits - lse[..., None])
    o = torch.einsum("hqk,kd->qhd", p, ckv.float())
    return o, (lse * LOG2E).transpose(0, 1).contiguous()


@dataclasses.dataclass(frozen=True)
class MLAHeadDimensions:
    """Dimensions of a single MLA head (reference mla/_core.py:76)."""
    qk_nope_head_dim: int
    qk_rope_head_dim: int
    v_head_dim: int
    kv_lora_rank: int


@dataclasses.dataclass(frozen=True)
class MLALayerDimensions:
    """Dimensions of an MLA layer (reference mla/_core.py:111)."""
    head_dimensions: MLAHeadDimensions
    num_heads: int


deepseek_mla_dimensions = MLAHeadDimensions(qk_nope_head_dim=128, qk_rope_head_dim=64, v_head_dim=128, kv_lora_rank=512)
smaller_mla_dimensions = MLAHeadDimensions(qk_nope_head_dim=64, qk_rope_head_dim=32, v_head_dim=64, kv_lora_rank=256)
supported_mla_layer_dimensions = [MLALayerDimensions(deepseek_mla_dimensions, 128), MLALayerDimensions(deepseek_mla_dimensions, 64)]


class BatchMLAPagedAttentionWrapper:
    """plan()/run() wrapper for MLA over a paged latent cache (``ckv [pages, page, 512]``, ``kpe [pages, page, 64]``)."""

    def __init__(self, float_workspace_buffer: torch.Tensor, use_cuda_graph: bool = False,
                 qo_indptr: Optional[torch.Tensor] = None, kv_indptr: Optional[torch.Tensor] = None,
                 kv_indices: Optional[torch.Tensor] = None, kv_len_arr: Optional[torch.Tensor] = None,
                 backend: str = "auto") -> None:
        self._float_workspace_buffer = float_workspace_buffer
        self.device = float_workspace_buffer.device
        self._use_cuda_graph = use_cuda_graph
        self._int_workspace_buffer = torch.empty(8 * 1024 * 1024, dtype=torch.uint8, device=self.device)
        self._pin_int_workspace_buffer = torch.empty(8 * 1024 * 1024, dtype=torch.uint8, device="cpu",
                                                     pin_memory=self.device.type == "cuda")
        self._planned = False

    def plan(self, qo_indptr: torch.Tensor, kv_indptr: torch.Tensor, kv_indices: torch.Tensor,
             kv_len_arr: torch.Tensor, num_heads: int, head_dim_ckv: int, head_dim_kpe: int, page_size: int,
             causal: bool, sm_scale: float, q_data_type: torch.dtype, kv_data_type: torch.dtype,
             use_profiler: bool = False) -> None:
        if head_dim_ckv != 512 or head_dim_kpe != 64:
            raise NotImplementedError("mla_sm100: only ckv=512 / kpe=64 is specialised")
        if num_heads > 128:
            raise NotImplementedError("mla_sm100: num_heads must be <= 128")
        self._num_heads, self._page_size = num_heads, page_size
        self._causal, self._sm_scale = bool(causal), float(sm_scale)
        self._q_dtype = q_data_type
        qo = qo_indptr.to("cpu", torch.int64)
        kvp = kv_indptr.to("cpu", torch.int64)
        kvl = kv_len_arr.to("cpu", torch.int64)
        self._qo_host, self._kvp_host, self._kvl_host = qo, kvp, kvl
        self._kv_indices = kv_indices.to(self.device, torch.int32)
        batch = kvl.numel()
        # ---- C++ planner (csrc/runtime/planner.cpp mla_plan): query tokens are flattened (q_len > 1: MTP / speculative decode, causal
        #      inside the new tokens), the KV chunk is the smallest multiple of the tile for which all (token, chunk) items fit one wave
        self._n_q = int(qo[-1]) if batch else 0
        ctas = max(1, device_sm_count(self.device if self.device.type == "cuda" else None) // 2)
        max_work = max(ctas, self._n_q, 1)
        pin = self._pin_int_workspace_buffer.view(torch.int32)
        if max_work * 8 + max(self._n_q, 1) > pin.numel():
            raise RuntimeError("int workspace too small for this MLA batch")
        work_h = pin[: max_work * 8]
        parts_h = pin[max_work * 8: max_work * 8 + max(self._n_q, 1)]
        counts = torch.zeros(4, dtype=torch.int64)
        jit.load("planner").call("mla_plan", qo.contiguous(), kvp.contiguous(), kvl.contiguous(), batch, 1 if causal else 0, ctas, _TILE,
                                 work_h, max_work, parts_h, max(self._n_q, 1), counts)
        self._num_work, kmax = int(counts[0]), int(counts[1])
        self._kmax = kmax
        n_used = max_work * 8 + max(self._n_q, 1)
        dev = self._int_workspace_buffer.view(torch.int32)
        dev[:n_used].copy_(pin[:n_used], non_blocking=self.device.type == "cuda")
        self._work = dev[: self._num_work * 8]
        self._row_parts = dev[max_work * 8: n_used]
        if kmax > 1:
            need = self._n_q * kmax * num_heads * (512 + 2) * 4
            if need > self._float_workspace_buffer.numel() * self._float_workspace_buffer.element_size():
                raise RuntimeError("float workspace too small for MLA split-KV partials")
            f = self._float_workspace_buffer.view(torch.uint8)[:need].view(torch.float32)
            self._partial_o = f[: self._n_q * kmax * num_heads * 512].view(self._n_q, kmax, num_heads, 512)
            self._partial_lse = f[self._n_q * kmax * num_heads * 512 : self._n_q * kmax * num_heads * 513].view(
                self._n_q, kmax, num_heads)
        self._planned = True

    def run(self, q_nope: torch.Tensor, q_pe: torch.Tensor, ckv_cache: torch.Tensor, kpe_cache: torch.Tensor,
            out: Optional[torch.Tensor] = None, lse: Optional[torch.Tensor] = None, return_lse: bool = False,
            profiler_buffer=None, kv_len=None, page_table=None, return_lse_base_on_e: bool = False,
            o_scale: Optional[float] = None):
        if not self._planned:
            raise RuntimeError("plan() must be called before run()")
        n, h = q_nope.shape[0], self._num_heads
        if out is None:
            out = torch.empty(n, h, 512, dtype=q_nope.dtype, device=q_nope.device)
        if return_lse and lse is None:
            lse = torch.empty(n, h, dtype=torch.float32, device=q_nope.device)
        if not q_nope.is_cuda:
            ps = self._page_size
            idx = self._kv_indices.cpu().long()
            for b in range(self._kvl_host.numel()):
                qs, qe = int(self._qo_host[b]), int(self._qo_host[b + 1])
                if qe == qs:
                    continue
                pages = idx[int(self._kvp_host[b]) : int(self._kvp_host[b + 1])]
                kvl = int(self._kvl_host[b])
                ckv = ckv_cache[pages].reshape(-1, 512)[:kvl]
                kpe = kpe_cache[pages].reshape(-1, 64)[:kvl]
                o, l = mla_attention_ref(q_nope[qs:qe], q_pe[qs:qe], ckv, kpe, self._sm_scale,
                                         kvl - (qe - qs) if self._causal else None)
                out[qs:qe] = o.to(out.dtype)
                if return_lse:
                    lse[qs:qe] = l
        else:
            if q_nope.dtype not in (torch.float16, torch.bfloat16) or ckv_cache.dtype != q_nope.dtype:
                raise NotImplementedError("mla_sm100: q/kv dtype must both be f16 or bf16")
            split = self._kmax > 1
            jit.load("mla_sm100").call(
                "mla_decode_run", q_nope, q_pe, ckv_cache, kpe_cache, self._kv_indices, self._work, self._num_work,
                out, self._partial_o if split else None, self._partial_lse if split else None,
                self._row_parts if split else None, self._kmax,
                lse if return_lse else None, n, h, self._page_size, ckv_cache.shape[0],
                q_nope.stride(0), q_nope.stride(1), q_pe.stride(0), q_pe.stride(1), ckv_cache.stride(0),
                ckv_cache.stride(1), kpe_cache.stride(0), kpe_cache.stride(1), out.stride(0), out.stride(1),
                self._sm_scale, dtype_code(q_nope.dtype), 1, stream_ptr(q_nope),
            )
        if o_scale is not None:
            out.mul_(o_scale)
        if return_lse and return_lse_base_on_e:
            lse = lse / LOG2E
        return (out, lse) if return_lse else out


def trtllm_batch_decode_with_kv_cache_mla(query: torch.Tensor, kv_cache: torch.Tensor, workspace_buffer: torch.Tensor,
                                          qk_nope_head_dim: int, kv_lora_rank: int, qk_rope_head_dim: int,
                                          block_tables: torch.Tensor, seq_lens: torch.Tensor, max_seq_len: int,
                                          sparse_mla_top_k: int = 0, out: Optional[torch.Tensor] = None,
                                          bmm1_scale: Union[float, torch.Tensor] = 1.0,
                                          bmm2_scale: Union[float, torch.Tensor] = 1.0, sinks=None,
                                          skip_softmax_threshold_scale_factor=None, enable_pdl=None, backend: str = "auto",
                                          is_var_seq: bool = True, uses_shared_paged_kv_idx: bool = True,
                                          lse: Optional[torch.Tensor] = None, return_lse: bool = False):
    """Function-style MLA decode (reference :631): ``query [B, q_len, H, 576]`` (nope|rope concatenated),
    ``kv_cache [pages, (1,) page, 576]``, ``block_tables [B, max_pages]``."""
    if sinks is not None:
        raise NotImplementedError("trtllm_batch_decode_with_kv_cache_mla: attention sinks are not implemented by the MLA kernel")
    if not uses_shared_paged_kv_idx:
        raise NotImplementedError("trtllm_batch_decode_with_kv_cache_mla: separate K / V page indices are not implemented")
    if sparse_mla_top_k:
        return _sparse_mla_decode(query, kv_cache, workspace_buffer, kv_lora_rank, qk_rope_head_dim, block_tables, sparse_mla_top_k,
                                  float(bmm1_scale), float(bmm2_scale), out, lse, return_lse)
    b, ql, h, _ = query.shape
    kvc = kv_cache.squeeze(1) if kv_cache.ndim == 4 else kv_cache
    page_size = kvc.shape[1]
    ckv, kpe = kvc[..., :kv_lora_rank], kvc[..., kv_lora_rank:]
    q = query.reshape(b * ql, h, -1)
    seq_host = seq_lens.to("cpu", torch.int64)
    npages = (seq_host + page_size - 1) // page_size
    kv_indptr = torch.zeros(b + 1, dtype=torch.int32)
    kv_indptr[1:] = npages.cumsum(0)
    bt = block_tables.to("cpu")
    kv_indices = torch.cat([bt[i, : int(npages[i])] for i in range(b)]).int() if b else torch.empty(0, dtype=torch.int32)
    w = BatchMLAPagedAttentionWrapper(workspace_buffer)
    scale = float(bmm1_scale)
    w.plan(torch.arange(0, (b + 1) * ql, ql, dtype=torch.int32), kv_indptr, kv_indices, seq_host.int(), h, kv_lora_rank,
           qk_rope_head_dim, page_size, True, scale, query.dtype, kv_cache.dtype)
    res = w.run(q[..., :kv_lora_rank], q[..., kv_lora_rank:], ckv, kpe, return_lse=return_lse, o_scale=float(bmm2_scale)
                if float(bmm2_scale) != 1.0 else None)
    o = res[0] if return_lse else res
    o = o.view(b, ql, h, kv_lora_rank)
    if out is not None:
        out.copy_(o)
        o = out
    return (o, res[1].view(b, ql, h)) if return_lse else o


def _sparse_mla_decode(query, kv_cache, workspace_buffer, kv_lora_rank, qk_rope_head_dim, block_tables, top_k, bmm1_scale,
                       bmm2_scale, out, lse, return_lse):
    """Sparse (top-k) MLA decode, DeepSeek-V3.2 style (reference flashinfer/mla/_core.py:631-940 ``sparse_mla_top_k``):
    ``block_tables [B, q_len, top_k]`` holds, for every query token, the row indices of its selected KV tokens in the flattened
    cache ``[pages * page_size, 576]`` (-1 = unused).  Two native kernels: ``gather_rows`` (csrc/elementwise/page.cu) packs the
    selected rows of every query token into a dense per-query cache, then the tcgen05 MLA kernel (csrc/attention/mla_sm100.cu)
    runs over it with full-size TMA boxes, every query token as its own request."""
    b, ql, h, dqk = query.shape
    if block_tables.shape != (b, ql, top_k):
        raise ValueError(f"Expected page_table.shape == (num_seqs, num_tokens, sparse_mla_top_k), got {tuple(block_tables.shape)}")
    kvc = kv_cache.squeeze(1) if kv_cache.ndim == 4 else kv_cache
    flat = kvc.reshape(-1, kvc.shape[-1])
    n = b * ql
    idx = block_tables.reshape(n, top_k).to(torch.int32)
    valid = idx >= 0
    lens = valid.sum(-1).to(torch.int32)
    if bool((valid[:, 1:] & ~valid[:, :-1]).any()):  # unused slots in the middle: move the valid indices to the front (stable)
        order = torch.argsort((~valid).to(torch.int8), dim=-1, stable=True)
        idx = torch.gather(idx, 1, order)
    page = 64
    kpad = (top_k + page - 1) // page * page
    if kpad != top_k:
        idx = torch.nn.functional.pad(idx, (0, kpad - top_k), value=-1)
    idx = idx.contiguous()
    dense = torch.empty(n * kpad, flat.shape[-1], dtype=flat.dtype, device=flat.device)
    if flat.is_cuda:
        esz = flat.element_size()
        if flat.stride(-1) != 1:
            flat = flat.contiguous()
        jit.load("page").call("gather_rows", flat, idx, dense, n * kpad, flat.shape[0], flat.shape[-1] * esz, flat.stride(0) * esz,
                              dense.stride(0) * esz, 1, stream_ptr(flat))
    else:
        dense.copy_(torch.where((idx >= 0).reshape(-1, 1), flat[idx.reshape(-1).clamp(min=0).long()], torch.zeros((), dtype=flat.dtype)))
    dense = dense.view(n * kpad // page, page, flat.shape[-1])
    ppr = kpad // page
    w = BatchMLAPagedAttentionWrapper(workspace_buffer)
    w.plan(torch.arange(0, n + 1, dtype=torch.int32), torch.arange(0, (n + 1) * ppr, ppr, dtype=torch.int32),
           torch.arange(0, n * ppr, dtype=torch.int32), lens.to("cpu"), h, kv_lora_rank, qk_rope_head_dim, page, False, bmm1_scale,
           query.dtype, dense.dtype)
    q = query.reshape(n, h, dqk)
    res = w.run(q[..., :kv_lora_rank], q[..., kv_lora_rank:], dense[..., :kv_lora_rank], dense[..., kv_lora_rank:],
                return_lse=return_lse, o_scale=bmm2_scale if bmm2_scale != 1.0 else None)
    o = (res[0] if return_lse else res).view(b, ql, h, kv_lora_rank)
    if out is not None:
        out.copy_(o)
        o = out
    if return_lse:
        l2 = res[1].view(b, ql, h)
        if lse is not None:
            lse.copy_(l2)
            l2 = lse
        return o, l2
    return o


def xqa_batch_decode_with_kv_cache_mla(query: torch.Tensor, kv_cache: torch.Tensor, workspace_buffer: torch.Tensor,
                                       qk_nope_head_dim: int, kv_lora_rank: int, qk_rope_head_dim: int, block_tables: torch.Tensor,
                                       seq_lens: torch.Tensor, max_seq_len: int, out: Optional[torch.Tensor] = None,
                                       bmm1_scale: Union[float, torch.Tensor] = 1.0, bmm2_scale: Union[float, torch.Tensor] = 1.0,
                                       sinks=None, enable_pdl=None):
    """sm120-only path in the reference (mla/_core.py:943); on B200 the tcgen05 MLA kernel serves the same API."""
    return trtllm_batch_decode_with_kv_cache_mla(query, kv_cache, workspace_buffer, qk_nope_head_dim, kv_lora_rank, qk_rope_head_dim,
                                                 block_tables, seq_lens, max_seq_len, out=out, bmm1_scale=bmm1_scale,
                                                 bmm2_scale=bmm2_scale, sinks=sinks, enable_pdl=enable_pdl)
