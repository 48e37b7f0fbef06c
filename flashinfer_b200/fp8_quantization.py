"""Module path of the reference (flashinfer/fp8_quantization.py); implementation: quantization/fp8.py."""
from .quantization.fp8 import mxfp8_dequantize_host, mxfp8_quantize  # noqa: F401
