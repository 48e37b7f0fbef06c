"""Paged KV-cache ops.  Parity: reference flashinfer/page.py:128-406."""
from __future__ import annotations

from typing import Optional, Tuple, Union

import torch

from . import jit
from .utils import dtype_code, get_seq_lens, paged_kv_strides, stream_ptr, unpack_paged_kv_cache  # noqa: F401


def get_batch_indices_positions(append_indptr: torch.Tensor, seq_lens: torch.Tensor, nnz: int,
                                batch_indices: Optional[torch.Tensor] = None,
                                positions: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """For each appended token: its request index and its absolute position in that request
    (``seq_lens`` are the lengths *after* the append)."""
    dev = append_indptr.device
    if batch_indices is None:
        batch_indices = torch.empty(nnz, dtype=torch.int32, device=dev)
    if positions is None:
        positions = torch.empty(nnz, dtype=torch.int32, device=dev)
    batch = seq_lens.numel()
    if not append_indptr.is_cuda:
        lens = (append_indptr[1:] - append_indptr[:-1]).long()
        b = torch.repeat_interleave(torch.arange(batch), lens)
        start = append_indptr[:-1].long()[b]
        pos = torch.arange(int(lens.sum())) - start + (seq_lens.long() - lens)[b]
        batch_indices[: b.numel()] = b.int()
        positions[: b.numel()] = pos.int()
        return batch_indices, positions
    jit.load("page").call("get_batch_indices_positions", append_indptr.int(), seq_lens.int(), batch_indices, positions,
                          batch, stream_ptr(append_indptr))
    return batch_indices, positions


def append_paged_kv_cache(append_key: torch.Tensor, append_value: torch.Tensor, batch_indices: torch.Tensor,
                          positions: torch.Tensor, paged_kv_cache: Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]],
                          kv_indices: torch.Tensor, kv_indptr: torch.Tensor, kv_last_page_len: torch.Tensor,
                          kv_layout: str = "NHD") -> None:
    """Scatter ragged K/V rows ``[nnz, H, D]`` into the paged cache."""
    k_cache, v_cache = unpack_paged_kv_cache(paged_kv_cache, kv_layout)
    sp, sn, sh, page_size, h, d = paged_kv_strides(k_cache, kv_layout)
    nnz = append_key.shape[0]
    if not append_key.is_cuda:
        pages = kv_indices[(kv_indptr[batch_indices.long()] + positions // page_size).long()].long()
        entry = (positions % page_size).long()
        if kv_layout == "NHD":
            k_cache[pages, entry] = append_key.to(k_cache.dtype)
            v_cache[pages, entry] = append_value.to(v_cache.dtype)
        else:
            k_cache[pages, :, entry] = append_key.to(k_cache.dtype)
            v_cache[pages, :, entry] = append_value.to(v_cache.dtype)
        return
    if append_key.dtype != k_cache.dtype:
        append_key, append_value = append_key.to(k_cache.dtype), append_value.to(v_cache.dtype)
    if paged_kv_strides(v_cache, kv_layout)[:3] != (sp, sn, sh):
        raise ValueError("k_cache and v_cache must share strides")
    jit.load("page").call(
        "append_paged_kv_cache", append_key, append_value, batch_indices, positions, k_cache, v_cache, kv_indices,
        kv_indptr, nnz, h, d, page_size, append_key.stride(0), append_key.stride(1), append_value.stride(0),
        append_value.stride(1), sp, sn, sh, dtype_code(k_cache.dtype), 1, stream_ptr(append_key),
    )


def append_paged_mla_kv_cache(append_ckv: torch.Tensor, append_kpe: torch.Tensor, batch_indices: torch.Tensor,
                              positions: torch.Tensor, ckv_cache: Optional[torch.Tensor],
                              kpe_cache: Optional[torch.Tensor], kv_indices: torch.Tensor, kv_indptr: torch.Tensor,
                              kv_last_page_len: torch.Tensor) -> None:
    """MLA latent cache append: ckv ``[nnz, 512]`` / kpe ``[nnz, 64]`` into ``[pages, page, dim]`` caches."""
    page_size = ckv_cache.shape[1]
    nnz = append_ckv.shape[0]
    if not append_ckv.is_cuda:
        pages = kv_indices[(kv_indptr[batch_indices.long()] + positions // page_size).long()].long()
        entry = (positions % page_size).long()
        ckv_cache[pages, entry] = append_ckv.to(ckv_cache.dtype)
        kpe_cache[pages, entry] = append_kpe.to(kpe_cache.dtype)
        return
    jit.load("page").call(
        "append_paged_mla_kv_cache", append_ckv, append_kpe, batch_indices, positions, ckv_cache, kpe_cache, kv_indices,
        kv_indptr, nnz, append_ckv.shape[-1], append_kpe.shape[-1], page_size, append_ckv.stride(0),
        append_kpe.stride(0), ckv_cache.stride(0), ckv_cache.stride(1), kpe_cache.stride(0), kpe_cache.stride(1),
        dtype_code(ckv_cache.dtype), 1, stream_ptr(append_ckv),
    )


def get_page_module(*args, **kwargs):
    """The native module behind this file's ops (reference page.py get_page_module: the JIT module accessor)."""
    from . import jit

    return jit.load("page")
