"""Module path of the reference (flashinfer/trtllm_low_latency_gemm.py); implementation: gemm/lowp.py."""
from .gemm import prepare_low_latency_gemm_weights, trtllm_low_latency_gemm  # noqa: F401
