"""DeepGEMM-style grouped fp8 GEMM entry points (reference flashinfer/deep_gemm.py: a JIT'd CUTLASS-free sm100 kernel there).
Here they are the grouped mode of the hand-written ``fp8_groupwise_kernel`` (csrc/gemm/gemm_blockscaled_sm100.cu)."""
from __future__ import annotations

import enum
from typing import Optional, Tuple

import torch

from .gemm.grouped import batch_deepgemm_fp8_nt_groupwise, group_deepgemm_fp8_nt_groupwise


class GemmType(enum.Enum):
    Normal = 0
    GroupedContiguous = 1
    GroupedMasked = 2


class MajorTypeAB(enum.Enum):
    KMajor = 0
    MNMajor = 1


class MajorTypeCD(enum.Enum):
    NMajor = 0
    MMajor = 1


def get_m_alignment_for_contiguous_layout() -> int:
    return 128


def get_device_arch() -> str:
    return "100a"


def must_be_k_major() -> bool:
    return True


def get_tma_aligned_size(x: int, element_size: int) -> int:
    align = 16 // element_size
    return (x + align - 1) // align * align


def m_grouped_fp8_gemm_nt_contiguous(a_fp8: Tuple[torch.Tensor, torch.Tensor], b_fp8: Tuple[torch.Tensor, torch.Tensor],
                                     d: torch.Tensor, m_indices: torch.Tensor, recipe=None, compiled_dims: str = "nk") -> None:
    """``d[i] = a[i] @ b[m_indices[i]].T`` with 1x128 / 128x128 fp32 scales; rows grouped per expert in 128-row blocks."""
    (a, sfa), (b, sfb) = a_fp8, b_fp8
    group_deepgemm_fp8_nt_groupwise(a, b, sfa, sfb, m_indices, out=d, out_dtype=d.dtype)


def m_grouped_fp8_gemm_nt_masked(a_fp8: Tuple[torch.Tensor, torch.Tensor], b_fp8: Tuple[torch.Tensor, torch.Tensor], d: torch.Tensor,
                                 masked_m: torch.Tensor, expected_m: int, recipe=None, compiled_dims: str = "nk") -> None:
    (a, sfa), (b, sfb) = a_fp8, b_fp8
    batch_deepgemm_fp8_nt_groupwise(a, b, sfa, sfb, masked_m, expected_m, out=d, out_dtype=d.dtype)


m_grouped_fp8_gemm_nt_contiguous_sm10x = m_grouped_fp8_gemm_nt_contiguous
m_grouped_fp8_gemm_nt_masked_sm10x = m_grouped_fp8_gemm_nt_masked
