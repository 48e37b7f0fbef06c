"""Module path of the reference (flashinfer/comm/trtllm_ar.py): fused all-reduce entry points (implementation: compat.py, allreduce.py)."""
from .compat import (  # noqa: F401
    AllReduceFusionOp,
    AllReduceFusionPattern,
    AllReduceStrategyConfig,
    AllReduceStrategyType,
    QuantizationSFLayout,
    compute_fp4_swizzled_layout_sf_size,
    trtllm_allreduce_fusion,
    trtllm_create_ipc_workspace_for_all_reduce_fusion,
    trtllm_custom_all_reduce,
    trtllm_destroy_ipc_workspace_for_all_reduce_fusion,
    trtllm_lamport_initialize,
    trtllm_lamport_initialize_all,
    trtllm_moe_allreduce_fusion,
    trtllm_moe_finalize_allreduce_fusion,
)
from . import trtllm_create_ipc_workspace_for_all_reduce, trtllm_destroy_ipc_workspace_for_all_reduce  # noqa: F401,E402
