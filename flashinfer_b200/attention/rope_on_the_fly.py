"""``pos_encoding_mode="ROPE_LLAMA"`` of the attention APIs (reference: RoPE applied to q and to the UN-rotated keys inside the
attention kernel, include/flashinfer/attention/prefill.cuh / decode.cuh with PosEncodingMode::kRoPELlama).

Here the mode is served by composition - rotate, then run the kernel in its plain mode: q is rotated with the native RoPE kernel at its
positions (queries are the last tokens of their request), the keys a call touches are rotated into a scratch copy (ragged keys: one
tensor; paged keys: the pages of the batch inside a clone of the cache, so page indices stay valid), values are used as they are.  It
costs one extra pass over the keys per call; engines that cache rotated keys (the normal case) should keep ``NONE``."""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def rotate_rows(x: torch.Tensor, positions: torch.Tensor, rope_scale: float, rope_theta: float) -> torch.Tensor:
    """RoPE (rotate-half pairing, the Llama convention of this mode) on ``x [n, H, D]`` at ``positions [n]``."""
    from ..rope import apply_rope_pos_ids

    if x.shape[0] == 0:
        return x
    pos = positions.to(x.device, torch.int32)
    return apply_rope_pos_ids(x.contiguous(), x[:, :1].contiguous(), pos, rope_scale=rope_scale, rope_theta=rope_theta)[0]


def query_positions(qo_indptr_host: torch.Tensor, kv_lens_host: torch.Tensor) -> torch.Tensor:
    """Position of every packed query row: the rows of a request are its LAST ``q_len`` tokens."""
    qo = qo_indptr_host.to(torch.int64)
    q_lens = qo[1:] - qo[:-1]
    start = kv_lens_host.to(torch.int64) - q_lens
    rows = torch.arange(int(qo[-1]))
    req = torch.searchsorted(qo, rows, right=True) - 1
    return (start[req] + rows - qo[:-1][req]).to(torch.int32)


def ragged_key_positions(kv_indptr_host: torch.Tensor) -> torch.Tensor:
    kv = kv_indptr_host.to(torch.int64)
    rows = torch.arange(int(kv[-1]))
    req = torch.searchsorted(kv, rows, right=True) - 1
    return (rows - kv[:-1][req]).to(torch.int32)


def rotated_paged_keys(k_cache: torch.Tensor, kv_indices: torch.Tensor, kv_indptr_host: torch.Tensor, kv_layout: str, rope_scale: float,
                       rope_theta: float) -> torch.Tensor:
    """Clone of ``k_cache`` in which the pages listed in ``kv_indices`` hold rotated keys (slot ``s`` of the ``j``-th page of a request
    is position ``j * page_size + s``; slots past the request's length are rotated too, harmlessly)."""
    pages = kv_indices.to(torch.int64)
    if pages.numel() == 0:
        return k_cache
    page_size = k_cache.shape[1] if kv_layout == "NHD" else k_cache.shape[2]
    ip = kv_indptr_host.to(torch.int64)
    idx = torch.arange(int(ip[-1]))
    req = torch.searchsorted(ip, idx, right=True) - 1
    page_no = (idx - ip[:-1][req]).to(k_cache.device)                                  # index of the page inside its request
    pos = (page_no[:, None] * page_size + torch.arange(page_size, device=k_cache.device)[None, :]).reshape(-1)
    used = k_cache[pages.to(k_cache.device)]
    rows = used if kv_layout == "NHD" else used.transpose(1, 2)                        # [n_pages, page_size, Hkv, D]
    shape = rows.shape
    rot = rotate_rows(rows.reshape(-1, shape[2], shape[3]), pos, rope_scale, rope_theta).reshape(shape)
    out = k_cache.clone()
    out[pages.to(k_cache.device)] = (rot if kv_layout == "NHD" else rot.transpose(1, 2)).to(k_cache.dtype)
    return out


def rope_params(rope_scale: Optional[float], rope_theta: Optional[float]) -> Tuple[float, float]:
    return (1.0 if rope_scale is None else float(rope_scale)), (1e4 if rope_theta is None else float(rope_theta))
