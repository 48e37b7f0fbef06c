"""Module path of the reference (flashinfer/comm/cuda_ipc.py)."""
from .compat import CudaRTLibrary, create_shared_buffer, free_shared_buffer  # noqa: F401
