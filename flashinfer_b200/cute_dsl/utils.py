"""Helpers of the reference's CuTe-DSL package that do not depend on the DSL (flashinfer/cute_dsl/utils.py): SM count, and the
views between the 128x4-swizzled scale-factor buffer and the 6-d "MMA layout" its grouped block-scaled GEMM consumes.  The kernels
of this library read the swizzled buffer directly (UTCCP of 32 x 16-byte chunks), so the 6-d form is only a view for callers that
build their tensors the reference's way."""
from typing import Tuple

import torch


def ceil_div(a: int, b: int) -> int:
    return -(-a // b)


def is_cute_dsl_available() -> bool:
    """The reference asks whether nvidia-cutlass-dsl can be imported before exposing the ops of this package; here those ops are
    native sm_100a kernels built from csrc/, so they are always present."""
    return True


def get_num_sm(device: torch.device) -> int:
    from ..utils import device_sm_count

    return device_sm_count(device)


def get_mma_sf_shape(m: int, k: int, num_groups: int = 1, sf_vec_size: int = 16) -> Tuple[int, int, int, int, int, int]:
    """Logical 6-d shape ``(row % 32, (row % 128) // 32, row_tile, col % 4, col_tile, group)`` of the scale factors of an ``[m, k]``
    operand: 128-row x 4-column tiles of one byte per ``sf_vec_size`` elements."""
    return (32, 4, ceil_div(m, 128), 4, ceil_div(ceil_div(k, sf_vec_size), 4), num_groups)


def convert_sf_to_mma_layout(sf: torch.Tensor, m: int, k: int, num_groups: int = 1, sf_vec_size: int = 16) -> torch.Tensor:
    """View a 128x4-swizzled scale-factor buffer (``fp4_quantize(..., is_sf_swizzled_layout=True)``, groups stacked along the rows)
    as the 6-d layout of :func:`get_mma_sf_shape`.  No bytes move: the swizzle stores, per (group, row tile, column tile), a
    ``[32][4][4]`` block indexed (row % 32, (row % 128) // 32, col % 4) - the result is that storage with its axes reordered, i.e. a
    non-contiguous view."""
    _, _, m_tiles, _, k_tiles, _ = get_mma_sf_shape(m, k, num_groups, sf_vec_size)
    want = num_groups * m_tiles * k_tiles * 512
    if sf.numel() != want:
        raise ValueError(f"scale-factor buffer has {sf.numel()} elements, a swizzled [{num_groups} x {m}, {k}] operand needs {want}")
    stored = sf.reshape(num_groups, m_tiles, k_tiles, 32, 4, 4)
    return stored.permute(3, 4, 1, 5, 2, 0)


def convert_sf_from_mma_layout(sf_6d: torch.Tensor, m: int, k: int, num_groups: int = 1, sf_vec_size: int = 16) -> torch.Tensor:
    """Inverse of :func:`convert_sf_to_mma_layout`: back to the 2-d swizzled buffer ``[num_groups * round_up(m, 128), round_up(k / vec, 4)]``
    (a copy when the input is the strided view)."""
    _, _, m_tiles, _, k_tiles, _ = get_mma_sf_shape(m, k, num_groups, sf_vec_size)
    stored = sf_6d.permute(5, 2, 4, 0, 1, 3).contiguous()
    return stored.reshape(num_groups * m_tiles * 128, k_tiles * 4)
