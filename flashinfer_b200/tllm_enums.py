"""Enumerations shared by the MoE / GEMM entry points (reference flashinfer/tllm_enums.py).  The integer values are part of
the public interface (callers pass them as plain ints), so they match the reference."""
from __future__ import annotations

from enum import IntEnum
from typing import Optional

import torch

from .fused_moe.core import ActivationType, Fp8QuantizationType, GatedActType, RoutingMethodType, WeightLayout  # noqa: F401
from .quantization import SfLayout  # noqa: F401


def _code(block_format: int, signed: int, integer: int, bits: int, uid: int) -> int:
    return (block_format << 24) | (signed << 20) | (integer << 16) | (bits << 8) | uid


class DtypeTrtllmGen(IntEnum):
    """Packed dtype descriptor: ``block-format << 24 | signed << 20 | integer << 16 | bits << 8 | uid``."""
    Bfloat16 = _code(0, 1, 0, 16, 0)
    Bool = _code(0, 0, 1, 1, 1)
    E2m1 = _code(1, 1, 0, 4, 2)
    E2m3 = _code(1, 1, 0, 6, 3)
    E3m2 = _code(1, 1, 0, 6, 4)
    E4m3 = _code(0, 1, 0, 8, 5)
    E5m2 = _code(0, 1, 0, 8, 6)
    Fp16 = _code(0, 1, 0, 16, 7)
    Fp32 = _code(0, 1, 0, 32, 8)
    Int8 = _code(0, 1, 1, 8, 9)
    Int32 = _code(0, 1, 1, 32, 10)
    Int64 = _code(0, 1, 1, 64, 11)
    MxE2m1 = _code(1, 1, 0, 4, 12)
    MxE4m3 = _code(1, 1, 0, 8, 13)
    MxInt4 = _code(1, 1, 1, 4, 14)
    UE8m0 = _code(0, 0, 0, 8, 15)
    UInt8 = _code(0, 0, 1, 8, 16)
    UInt16 = _code(0, 0, 1, 16, 17)
    UInt32 = _code(0, 0, 1, 32, 18)
    UInt64 = _code(0, 0, 1, 64, 19)
    UInt128 = _code(0, 0, 1, 128, 20)
    Void = _code(0, 1, 0, 0, 21)


def trtllm_gen_dtype_has_scale(dtype: DtypeTrtllmGen) -> bool:
    return dtype in (DtypeTrtllmGen.E2m1, DtypeTrtllmGen.MxE2m1, DtypeTrtllmGen.MxE4m3, DtypeTrtllmGen.MxInt4)


def deduce_trtllm_gen_tensor_dtype(x: torch.Tensor, scale: Optional[torch.Tensor]) -> DtypeTrtllmGen:
    if x.dtype == torch.bfloat16:
        return DtypeTrtllmGen.Bfloat16
    if x.dtype == torch.float16:
        return DtypeTrtllmGen.Fp16
    if x.dtype == torch.float8_e4m3fn:
        return DtypeTrtllmGen.E4m3 if scale is None else DtypeTrtllmGen.MxE4m3
    if x.dtype == torch.uint8:
        if scale is None:
            raise ValueError("Scale tensor must be provided for packed fp4 input")
        return DtypeTrtllmGen.E2m1 if scale.numel() == x.numel() * 2 // 16 else DtypeTrtllmGen.MxE2m1
    raise ValueError("Unsupported trtllm-gen input tensor.")
