"""flashinfer_b200 — a B200-native (sm_100a) LLM-inference kernel library with FlashInfer's API surface."""
__version__ = "0.1.0"

from . import jit, reference, utils  # noqa: F401
from .decode import (  # noqa: F401
    BatchDecodeWithPagedKVCacheWrapper,
    CUDAGraphBatchDecodeWithPagedKVCacheWrapper,
    fast_decode_plan,
    single_decode_with_kv_cache,
)
from .gemm import bmm_bf16, mm_bf16, tgv_gemm_sm100  # noqa: F401
from .utils import MaskMode, PosEncodingMode, TensorLayout, next_positive_power_of_2  # noqa: F401
