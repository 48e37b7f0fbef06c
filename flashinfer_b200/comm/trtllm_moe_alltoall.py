"""Module path of the reference (flashinfer/comm/trtllm_moe_alltoall.py): expert-parallel dispatch / combine (implementation: moe_alltoall.py)."""
from .moe_alltoall import (  # noqa: F401
    MoeAlltoAll,
    moe_a2a_combine,
    moe_a2a_dispatch,
    moe_a2a_get_workspace_size_per_rank,
    moe_a2a_initialize,
    moe_a2a_sanitize_expert_ids,
    moe_a2a_wrap_payload_tensor_in_workspace,
)
