"""Module path of the reference (flashinfer/fp4_quantization.py); implementation: quantization/fp4.py."""
from .quantization import *  # noqa: F401,F403
from .quantization import get_fp4_quantization_module  # noqa: F401
from .quantization.fp4 import (  # noqa: F401
    nvfp4_dequantize_paged_kv_cache,
    nvfp4_kv_dequantize,
    nvfp4_kv_quantize,
    nvfp4_quantize_paged_kv_cache,
    scaled_fp4_grouped_quantize,
)
