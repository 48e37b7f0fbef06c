"""Shared enums, dtype codes, layout helpers and argument checks.

Parity: reference flashinfer/utils.py:32-51 (PosEncodingMode / MaskMode / TensorLayout),
``_unpack_paged_kv_cache`` (flashinfer/utils.py), ``get_seq_lens`` (flashinfer/page.py:203-228).
"""
from __future__ import annotations

import math
from enum import Enum
from typing import Optional, Tuple, Union

import torch


class PosEncodingMode(Enum):
    NONE = 0
    ROPE_LLAMA = 1
    ALIBI = 2


class MaskMode(Enum):
    NON_CAUSAL = 0
    CAUSAL = 1
    CUSTOM = 2
    MULTIITEMSCORING = 3


class TensorLayout(Enum):
    NHD = 0
    HND = 1


# dtype codes shared with csrc/include/fib200/common.cuh
DTYPE_CODE = {
    torch.float16: 0,
    torch.bfloat16: 1,
    torch.float32: 2,
    torch.float8_e4m3fn: 3,
    torch.float8_e5m2: 4,
    torch.uint8: 5,
    torch.int32: 6,
    torch.int64: 7,
}


def dtype_code(dt: torch.dtype) -> int:
    try:
        return DTYPE_CODE[dt]
    except KeyError as e:
        raise TypeError(f"unsupported dtype {dt}") from e


def check_kv_layout(kv_layout: str) -> None:
    if kv_layout not in ("NHD", "HND"):
        raise KeyError(f"Invalid kv_layout {kv_layout}")


def check_pos_encoding_mode(mode: str) -> None:
    if not hasattr(PosEncodingMode, mode):
        raise KeyError(f"Invalid pos_encoding_mode {mode}")


def is_cuda(*tensors) -> bool:
    return any(isinstance(t, torch.Tensor) and t.is_cuda for t in tensors)


def next_positive_power_of_2(x: int) -> int:
    if x < 1:
        return 1
    return 1 << (x - 1).bit_length()


def ceil_div(a: int, b: int) -> int:
    return (a + b - 1) // b


def round_up(a: int, b: int) -> int:
    return ceil_div(a, b) * b


def unpack_paged_kv_cache(
    paged_kv_cache: Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]], kv_layout: str
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Accept ``(k_cache, v_cache)`` or a combined 5-D ``[pages, 2, ...]`` tensor and return
    4-D views ``[pages, page, H, D]`` (NHD) / ``[pages, H, page, D]`` (HND) without copying."""
    if isinstance(paged_kv_cache, tuple):
        k, v = paged_kv_cache
    elif torch.is_tensor(paged_kv_cache):
        if paged_kv_cache.ndim != 5:
            raise ValueError(f"combined paged_kv_cache must be 5-D, got {paged_kv_cache.ndim}-D")
        k, v = paged_kv_cache[:, 0], paged_kv_cache[:, 1]
    else:
        raise TypeError("paged_kv_cache must be a tensor or a (k_cache, v_cache) tuple")
    if k.ndim == 3:  # page_size 1 shorthand [pages, H, D]
        k = k.unsqueeze(1 if kv_layout == "NHD" else 2)
        v = v.unsqueeze(1 if kv_layout == "NHD" else 2)
    return k, v


def paged_kv_strides(k_cache: torch.Tensor, kv_layout: str):
    """(stride_page, stride_n, stride_h, page_size, num_kv_heads, head_dim) in elements."""
    if kv_layout == "NHD":
        pages, page_size, h, d = k_cache.shape
        sp, sn, sh, sd = k_cache.stride()
    else:
        pages, h, page_size, d = k_cache.shape
        sp, sh, sn, sd = k_cache.stride()
    if sd != 1:
        raise ValueError("last dim of the KV cache must be contiguous")
    return sp, sn, sh, page_size, h, d


def get_seq_lens(kv_indptr: torch.Tensor, kv_last_page_len: torch.Tensor, page_size: int) -> torch.Tensor:
    """kv_len = (n_pages - 1) * page_size + last_page_len (0 pages -> 0)."""
    n_pages = kv_indptr[1:] - kv_indptr[:-1]
    return torch.clamp(n_pages - 1, min=0) * page_size + torch.where(
        n_pages > 0, kv_last_page_len, torch.zeros_like(kv_last_page_len)
    )


def default_sm_scale(head_dim: int) -> float:
    return 1.0 / math.sqrt(head_dim)


LOG2E = 1.4426950408889634


def device_sm_count(device=None) -> int:
    if torch.cuda.is_available():
        return torch.cuda.get_device_properties(device or torch.cuda.current_device()).multi_processor_count
    return 148


def is_sm100a_supported(device=None) -> bool:
    if not torch.cuda.is_available():
        return False
    major, _ = torch.cuda.get_device_capability(device)
    return major == 10


def stream_ptr(t: Optional[torch.Tensor] = None) -> int:
    dev = t.device if t is not None else None
    return torch.cuda.current_stream(dev).cuda_stream


def device_support_pdl(device=None) -> bool:
    if not torch.cuda.is_available():
        return False
    major, _ = torch.cuda.get_device_capability(device)
    return major >= 9


def tensor_stats(t: torch.Tensor) -> dict:
    """min / max / mean / NaN count / Inf count in one pass (native kernel on CUDA; used by API logging level 5)."""
    n = t.numel()
    if not t.is_cuda or t.dtype not in (torch.float16, torch.bfloat16, torch.float32):
        f = t.detach().float().flatten()
        fin = f[torch.isfinite(f)]
        return {"min": float(fin.min()) if fin.numel() else float("nan"), "max": float(fin.max()) if fin.numel() else float("nan"),
                "mean": float(fin.mean()) if fin.numel() else float("nan"), "nan": int(torch.isnan(f).sum()),
                "inf": int(torch.isinf(f).sum())}
    from .. import jit

    out = torch.tensor([float("inf"), float("-inf"), 0.0, 0.0, 0.0, float("inf")], dtype=torch.float32, device=t.device)
    x = t.contiguous()
    jit.load("runtime").call("tensor_stats", x, n, dtype_code(x.dtype), out, stream_ptr(x))
    o = out.tolist()
    finite = max(n - int(o[3]) - int(o[4]), 1)
    return {"min": o[5], "max": o[1], "mean": o[2] / finite, "nan": int(o[3]), "inf": int(o[4])}
