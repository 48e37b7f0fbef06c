"""Module path of the reference (flashinfer/comm/dlpack_utils.py)."""
from .compat import pack_strided_memory  # noqa: F401
