"""Radix top-k and the fused top-k -> page-table / ragged-index transforms.
Parity: reference flashinfer/topk.py:508-911."""
from __future__ import annotations

from enum import IntEnum
from typing import Optional, Tuple

import torch

from . import jit
from .utils import dtype_code, stream_ptr


class TopKTieBreak(IntEnum):
    NONE = 0
    SMALL = 1
    LARGE = 2

    def __str__(self) -> str:
        return self.name.lower()


def _run(input, k, mode, lengths=None, row_starts=None, row_to_batch=None, page_table=None, ragged_offsets=None,
         tie_break=0, want_values=False, clusters=0, pdl=False):
    """``clusters``: 0 = auto (one row per thread-block cluster of 2 / 4 / 8 CTAs when the rows alone cannot fill the SMs),
    1 = one CTA per row, 2 / 4 / 8 = cluster size (csrc/elementwise/topk.cu: topk_kernel / topk_cluster_kernel)."""
    rows, max_len = input.shape
    x = input if input.stride(-1) == 1 else input.contiguous()
    idx = torch.empty(rows, k, dtype=torch.int32, device=x.device)
    vals = torch.empty(rows, k, dtype=x.dtype, device=x.device) if want_values else None
    i32 = lambda t: t.to(torch.int32).contiguous() if t is not None else None  # noqa: E731
    jit.load("topk").call(
        "topk_run_ex", x, x.stride(0), vals, idx, i32(lengths), i32(row_starts), i32(row_to_batch), i32(page_table),
        page_table.stride(0) if page_table is not None else 0, i32(ragged_offsets), rows, max_len, k, mode,
        int(tie_break), dtype_code(x.dtype), int(clusters), 1 if pdl else 0, stream_ptr(x),
    )
    return vals, idx


def top_k(input: torch.Tensor, k: int, sorted: bool = False, deterministic: bool = False,
          tie_break: int = TopKTieBreak.NONE, dsa_graph_safe: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k of every row of ``input [rows, n]``: returns ``(values [rows,k], indices [rows,k] int64)``."""
    if not input.is_cuda:
        v, i = torch.topk(input, k, dim=-1, sorted=sorted)
        return v, i
    vals, idx = _run(input, k, 0, tie_break=tie_break, want_values=True)
    if sorted:
        order = torch.argsort(vals.float(), dim=-1, descending=True, stable=True)
        vals, idx = vals.gather(-1, order), idx.gather(-1, order)
    return vals, idx.long()


def _transform_cpu(input, lengths, k, row_starts, tie_break):
    rows = input.shape[0]
    out = torch.full((rows, k), -1, dtype=torch.int64)
    for r in range(rows):
        s = int(row_starts[r]) if row_starts is not None else 0
        n = int(lengths[r])
        kk = min(k, n)
        seg = input[r, s : s + n].float()
        if tie_break == 2:
            order = torch.argsort(seg.flip(0), descending=True, stable=True)
            sel = (n - 1 - order[:kk])
        else:
            sel = torch.argsort(seg, descending=True, stable=True)[:kk]
        out[r, :kk] = torch.sort(sel).values if tie_break != 2 else torch.sort(sel, descending=True).values
    return out


def top_k_page_table_transform(input: torch.Tensor, src_page_table: torch.Tensor, lengths: torch.Tensor, k: int,
                               row_to_batch: Optional[torch.Tensor] = None, deterministic: bool = False,
                               tie_break: int = TopKTieBreak.NONE, dsa_graph_safe: bool = False,
                               row_starts: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``out[i, j] = src_page_table[batch(i), row_start(i) + topk_idx(i)[j]]`` (``-1`` padded), int32."""
    if not input.is_cuda:
        sel = _transform_cpu(input, lengths, k, row_starts, int(tie_break))
        out = torch.full(sel.shape, -1, dtype=torch.int32)
        for r in range(sel.shape[0]):
            b = int(row_to_batch[r]) if row_to_batch is not None else r
            s = int(row_starts[r]) if row_starts is not None else 0
            m = sel[r] >= 0
            out[r, m] = src_page_table[b, (sel[r, m] + s).long()].int()
        return out
    return _run(input, k, 1, lengths, row_starts, row_to_batch, src_page_table, tie_break=tie_break)[1]


def top_k_ragged_transform(input: torch.Tensor, offsets: torch.Tensor, lengths: torch.Tensor, k: int,
                           deterministic: bool = False, tie_break: int = TopKTieBreak.NONE,
                           dsa_graph_safe: bool = False, row_starts: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``out[i, j] = offsets[i] + row_start(i) + topk_idx(i)[j]`` (``-1`` padded), int32."""
    if not input.is_cuda:
        sel = _transform_cpu(input, lengths, k, row_starts, int(tie_break))
        s = row_starts.long() if row_starts is not None else torch.zeros(sel.shape[0], dtype=torch.long)
        out = torch.where(sel >= 0, sel + (offsets.long() + s)[:, None], torch.full_like(sel, -1))
        return out.int()
    return _run(input, k, 2, lengths, row_starts, None, None, offsets, tie_break=tie_break)[1]


# ------------------------------------------------------------------ reference helper names (flashinfer/topk.py:354-504)
def can_implement_filtered_topk() -> bool:
    """The reference's FilteredTopK needs 128 KB of dynamic shared memory; every B200 has 227 KB."""
    return True


def roundup_kbyte(x: int) -> int:
    return (x + 1023) // 1024 * 1024


def get_num_cached_for_topk(device, k: int) -> int:
    """Candidate-cache capacity of the reference's cluster top-k (kept for callers that size buffers with it)."""
    shared_per_block = (227 * 1024) // 2
    return (shared_per_block - (k + 5 + 3 * 256 + 8) * 4 - 1024) // 16


def get_fast_topk_clusters(batch_size: int) -> int:
    return 8 if batch_size <= 32 else (4 if batch_size < 128 else (2 if batch_size < 256 else 1))


def can_use_clusters_topk(device, deterministic: bool, dsa_graph_safe: bool) -> bool:
    return not dsa_graph_safe and not deterministic


def _cluster_size(rows: int) -> int:
    return max(2, get_fast_topk_clusters(rows))


def topk_clusters_exact(logits: torch.Tensor, top_k: int, output_values: bool = False, out_dtype=torch.int32, pdl: bool = False):
    """Exact top-k indices (and optionally values) per row on the cluster kernel: one row per thread-block cluster, slices of
    the row per CTA, histograms and compaction bases merged through distributed shared memory
    (reference include/flashinfer/fast_topk_clusters_exact.cuh:408-497)."""
    if not logits.is_cuda:
        vals, idx = torch.topk(logits, top_k, dim=-1, sorted=False)
        return (idx.to(out_dtype), vals) if output_values else idx.to(out_dtype)
    vals, idx = _run(logits, top_k, 0, want_values=output_values, clusters=_cluster_size(logits.shape[0]), pdl=pdl)
    idx = idx.to(out_dtype)
    return (idx, vals) if output_values else idx


def topk_clusters_page_table_transform(logits, seq_lens, src_page_table, top_k: int, pdl: bool = False):
    """Cluster flavour of :func:`top_k_page_table_transform` (rows = query tokens, one request each)."""
    if not logits.is_cuda:
        return top_k_page_table_transform(logits, src_page_table, seq_lens, top_k)
    return _run(logits, top_k, 1, lengths=seq_lens, page_table=src_page_table, clusters=_cluster_size(logits.shape[0]), pdl=pdl)[1]


def topk_clusters_ragged_transform(logits, seq_lens, offsets, top_k: int, pdl: bool = False):
    """Cluster flavour of :func:`top_k_ragged_transform`."""
    if not logits.is_cuda:
        return top_k_ragged_transform(logits, offsets, seq_lens, top_k)
    return _run(logits, top_k, 2, lengths=seq_lens, ragged_offsets=offsets, clusters=_cluster_size(logits.shape[0]), pdl=pdl)[1]


def get_topk_module(*args, **kwargs):
    """The native module behind this file's ops (reference topk.py get_topk_module: the JIT module accessor)."""
    from . import jit

    return jit.load("topk")
