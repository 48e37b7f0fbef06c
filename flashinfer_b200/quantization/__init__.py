"""Quantisation ops (NVFP4 / MXFP4 / MXFP8 / bit packing).  Parity: reference flashinfer/quantization/."""
from .fp4 import (  # noqa: F401
    SfLayout,
    block_scale_interleave,
    e2m1_and_ufp8sf_scale_to_float,
    fp4_quantize,
    mxfp4_dequantize,
    mxfp4_dequantize_host,
    mxfp4_quantize,
    nvfp4_batched_quantize,
    nvfp4_block_scale_interleave,
    nvfp4_kv_dequantize,
    nvfp4_kv_quantize,
    nvfp4_quantize,
    scaled_fp4_grouped_quantize,
    shuffle_matrix_a,
    shuffle_matrix_sf_a,
)
from .fp8 import mxfp8_dequantize_host, mxfp8_quantize  # noqa: F401
from .packbits import packbits, segment_packbits  # noqa: F401

# CuTe-DSL entry points of the reference resolve to the native kernels
nvfp4_quantize_cute_dsl = nvfp4_quantize
mxfp4_quantize_cute_dsl = mxfp4_quantize
mxfp8_quantize_cute_dsl = mxfp8_quantize


def is_cute_dsl_available() -> bool:
    return False  # nothing here depends on nvidia-cutlass-dsl


def get_fp4_quantization_module(backend: str = "100"):
    from .. import jit

    return jit.load("quantization")

from .. import _alias  # noqa: E402

_alias.install(__name__, ['fp4_quantization', 'fp8_quantization'])  # the reference's per-file module paths


from .. import jit as _jit_acc  # noqa: E402

get_mxfp8_quantization_sm100_module = _jit_acc.module_accessor("quantization")
get_fp4_kv_quantization_module = _jit_acc.module_accessor("quantization")
get_fp4_kv_dequantization_module = _jit_acc.module_accessor("quantization")
gen_fp4_quantization_sm100_module = _jit_acc.module_accessor("quantization")
