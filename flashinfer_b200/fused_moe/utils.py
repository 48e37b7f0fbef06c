"""Host-side helpers of the fused-MoE stack: token-count tuning buckets, swizzled scale-factor shapes and layout
conversion, compile-mode flags.

Parity: reference flashinfer/fused_moe/utils.py (bucket helpers :173-388, swizzle helpers :86-170, fp4 shapes :390-413,
flags :24-36 and :416-424).  The bucket grid is the reference's four-phase spacing (x2 up to 256, +256 up to 2048, +512 up
to 4096, x2 beyond) because tuned-config files are keyed by it; lookups here are bisects over the generated grid.
"""
from __future__ import annotations

import bisect
import contextlib
import threading
from dataclasses import dataclass
from typing import Callable, Dict, List, Sequence, Tuple

import torch

from ..utils import ceil_div, round_up

_flags = {"torch_compiling": False, "piecewise_cuda_graph": False}
_tls = threading.local()


def set_torch_compiling(enable: bool) -> None:
    _flags["torch_compiling"] = bool(enable)


def is_torch_compiling() -> bool:
    return _flags["torch_compiling"]


def set_piecewise_cuda_graph_flag(enable: bool) -> None:
    _flags["piecewise_cuda_graph"] = bool(enable)


def get_piecewise_cuda_graph_flag() -> bool:
    return _flags["piecewise_cuda_graph"]


def get_global_attrs():
    return _tls


def get_model_extra_attrs():
    return getattr(_tls, "attrs", None)


@contextlib.contextmanager
def model_extra_attrs(attrs: Dict):
    """Thread-local side channel for per-model attributes visible to ops called underneath."""
    prev = getattr(_tls, "attrs", None)
    _tls.attrs = attrs
    try:
        yield
    finally:
        _tls.attrs = prev


def with_model_extra_attrs(get_attrs: Callable):
    def deco(fn):
        def wrapped(self, *args, **kwargs):
            with model_extra_attrs(get_attrs(self)):
                return fn(self, *args, **kwargs)

        wrapped.__name__, wrapped.__doc__ = fn.__name__, fn.__doc__
        return wrapped

    return deco


# ---------------------------------------------------------------- scale-factor layouts
@dataclass
class Fp4QuantizedTensor:
    fp4_tensor: torch.Tensor
    scaling_factor: torch.Tensor
    is_sf_swizzled: bool = True

    @property
    def shape(self):
        return self.fp4_tensor.shape


def compute_swizzled_sf_shape(row: int, col: int) -> Tuple[int, int]:
    """Padded (rows, scale columns) of the 128x4-tiled layout."""
    return round_up(row, 128), round_up(col, 4)


def swizzle_sf(sf: torch.Tensor, rows: int, cols: int, scaling_vector_size: int = 16) -> torch.Tensor:
    """Row-major ``[b,] rows x ceil(cols / vec)`` scale factors -> flat 128x4-tiled layout (padded)."""
    from ..quantization.fp4 import block_scale_interleave

    kc = ceil_div(cols, scaling_vector_size)
    return block_scale_interleave(sf.reshape(-1, rows, kc)).reshape(-1)


def unswizzle_sf(sf: torch.Tensor, rows: int, cols: int, scaling_vector_size: int = 16) -> torch.Tensor:
    """Inverse of :func:`swizzle_sf`: flat tiled layout -> ``[b * rows, ceil(cols / vec)]``."""
    from ..quantization.fp4 import _swizzled_sf_size, _unswizzle_index

    kc = ceil_div(cols, scaling_vector_size)
    per = _swizzled_sf_size(rows, kc)
    flat = sf.reshape(-1).view(torch.uint8)
    if flat.numel() % per:
        raise ValueError(f"{flat.numel()} scale bytes is not a multiple of the padded tile size {per}")
    idx = _unswizzle_index(rows, kc).to(flat.device)
    return flat.view(-1, per)[:, idx].reshape(-1, kc).view(sf.dtype)


def reswizzle_sf(sf: torch.Tensor, rows: int, cols: int, scaling_vector_size: int = 16) -> torch.Tensor:
    """``sf`` holds several independently tiled partitions of ``rows`` rows each; return the tiling of their row-wise
    concatenation (what a GEMM over the stacked rows expects)."""
    kc = ceil_div(cols, scaling_vector_size)
    pr, pc = compute_swizzled_sf_shape(rows, kc)
    if sf.numel() % (pr * pc):
        raise ValueError("scale tensor is not a whole number of padded partitions")
    parts = sf.numel() // (pr * pc)
    linear = unswizzle_sf(sf, rows, cols, scaling_vector_size)            # [parts * rows, kc]
    return swizzle_sf(linear.reshape(parts * rows, kc), parts * rows, cols, scaling_vector_size)


def get_fp4_shape(input_shape: Sequence[int], sf_vec_size: int, is_swizzled_layout: bool = True):
    """(packed e2m1 shape, number of scale bytes) for a tensor of ``input_shape``."""
    m = 1
    for d in input_shape[:-1]:
        m *= d
    kc = input_shape[-1] // sf_vec_size
    packed = list(input_shape[:-1]) + [input_shape[-1] // 2]
    return packed, (round_up(m, 128) * round_up(kc, 4) if is_swizzled_layout else m * kc)


def fp4_scale_infer_shape(input_shapes: List[List[int]]) -> int:
    return get_fp4_shape(input_shapes[0], 16)[1]


# ---------------------------------------------------------------- tuning buckets
def next_positive_power_of_2(x: int) -> int:
    return 1 if x < 1 else 1 << (x - 1).bit_length()


def last_positive_power_of_2(x: int) -> int:
    return 1 if x < 1 else 1 << (x.bit_length() - 1)


def nearest_in_buckets(x: int, buckets: Sequence[int]) -> int:
    return min(max(next_positive_power_of_2(x), buckets[0]), buckets[-1])


def get_last_power_of_2_num_tokens_buckets(max_num_tokens: int, min_num_tokens: int = 1) -> Tuple[int, ...]:
    top = last_positive_power_of_2(max_num_tokens)
    out, m = [], max(1, next_positive_power_of_2(min_num_tokens))
    while m <= top:
        out.append(m)
        m *= 2
    return tuple(out) or (top,)


# (upper bound of the phase, step); step 0 = doubling
_PHASES = ((256, 0), (2048, 256), (4096, 512))


def _snap_up(x: int) -> int:
    for hi, step in _PHASES:
        if x <= hi:
            return next_positive_power_of_2(x) if step == 0 else ceil_div(x, step) * step
    return next_positive_power_of_2(x)


def get_hybrid_num_tokens_buckets(max_num_tokens: int, min_num_tokens: int = 1) -> Tuple[int, ...]:
    """Token-count grid with finer spacing where MoE tile counts change fastest; always contains ``max_num_tokens``."""
    grid = set()
    m = max(min_num_tokens, 1)
    while m <= min(max_num_tokens, 256):
        grid.add(m)
        m *= 2
    grid.update(range(512, min(max_num_tokens, 2048) + 1, 256))
    grid.update(range(2560, min(max_num_tokens, 4096) + 1, 512))
    m = 8192
    while m <= max_num_tokens:
        grid.add(m)
        m *= 2
    grid.add(max_num_tokens)
    return tuple(sorted(grid))


def map_to_hybrid_bucket(x: int, max_num_tokens: int) -> int:
    if x <= 0:
        return 1
    return max_num_tokens if x >= max_num_tokens else min(_snap_up(x), max_num_tokens)


def map_to_hybrid_bucket_uncapped(x: int) -> int:
    return 1 if x <= 0 else _snap_up(x)


def round_to_nearest_bucket(x: int, buckets: Sequence[int], round_map: bool = False) -> int:
    """Floor (default) or ceil of ``x`` onto ascending ``buckets``, clamped to their range."""
    if len(buckets) == 0:
        raise ValueError("buckets must be non-empty")
    if round_map:
        i = bisect.bisect_left(buckets, x)
        return buckets[min(i, len(buckets) - 1)]
    i = bisect.bisect_right(buckets, x) - 1
    return buckets[max(i, 0)]


def make_bucket_mapper(buckets: Tuple[int, ...], round_map: bool = False) -> Callable[[int], int]:
    if len(buckets) == 0:
        raise ValueError("buckets must be non-empty")
    grid = tuple(sorted(set(buckets)))
    return lambda x: round_to_nearest_bucket(x, grid, round_map)
