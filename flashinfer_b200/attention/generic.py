"""Python launcher for the catch-all CUDA-core attention kernel (csrc/attention/generic_attention.cu): head dims other
than 128, fp8 KV caches, custom masks, ALiBi.  The wrappers route here whenever the tcgen05 kernels do not cover a
configuration, so every configuration runs native code on the GPU."""
from __future__ import annotations

from typing import Optional

import torch

from .. import jit
from ..utils import dtype_code, stream_ptr

_KV_CODES = {torch.float16: 0, torch.bfloat16: 1, torch.float8_e4m3fn: 3, torch.float8_e5m2: 4}


def supported(q: torch.Tensor, k: torch.Tensor, dqk: int, dvo: int) -> bool:
    return (q.is_cuda and q.dtype in (torch.float16, torch.bfloat16) and k.dtype in _KV_CODES and dqk <= 256 and dvo <= 256
            and dqk % 16 == 0)


def pack_mask_bits(mask_bool: torch.Tensor) -> torch.Tensor:
    """flat bool -> little-endian packed uint8 (bit i of the stream = element i)."""
    m = mask_bool.flatten().to(torch.uint8)
    pad = (-m.numel()) % 8
    if pad:
        m = torch.cat([m, torch.zeros(pad, dtype=torch.uint8, device=m.device)])
    w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.uint8, device=m.device)
    return (m.view(-1, 8) * w).sum(1).to(torch.uint8)


def run(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, lse: Optional[torch.Tensor],
        qo_indptr: torch.Tensor, kv_indptr: torch.Tensor, kv_indices: Optional[torch.Tensor],
        kv_last_page_len: Optional[torch.Tensor], page_size: int, k_strides, v_strides, num_kv_heads: int, causal: bool,
        window_left: int, sm_scale: float, soft_cap: float = 0.0, packed_mask: Optional[torch.Tensor] = None,
        mask_indptr: Optional[torch.Tensor] = None, alibi_slopes: Optional[torch.Tensor] = None, k_scale: float = 1.0,
        v_scale: float = 1.0, enable_pdl: bool = True) -> None:
    """q ``[total_q, Hq, Dqk]``; ragged k/v ``[total_kv, Hkv, D]`` (``kv_indices=None``, strides = (0, sn, sh)) or paged
    caches with ``k_strides = (page, token, head)`` element strides."""
    total_q, hq, dqk = q.shape
    dvo = out.shape[-1]
    strides = torch.tensor([q.stride(0), q.stride(1), out.stride(0), out.stride(1), *k_strides, *v_strides, 0, 0, 0, 0, 0, 0],
                           dtype=torch.int64)
    jit.load("attention_generic").call(
        "generic_attention", q, k, v, out, lse, qo_indptr, kv_indptr, kv_indices, kv_last_page_len, packed_mask, mask_indptr,
        alibi_slopes, strides, qo_indptr.numel() - 1, total_q, hq, num_kv_heads, dqk, dvo, page_size, 1 if causal else 0,
        int(window_left), float(sm_scale), float(soft_cap), float(k_scale), float(v_scale), dtype_code(q.dtype),
        _KV_CODES[k.dtype], 1 if enable_pdl else 0, stream_ptr(q))
