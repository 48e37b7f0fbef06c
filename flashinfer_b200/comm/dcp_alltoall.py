"""Module path of the reference (flashinfer/comm/dcp_alltoall.py): decode context-parallel all-to-all (implementation: collectives.py)."""
from .collectives import (  # noqa: F401
    decode_cp_a2a_allocate_mnnvl_workspace,
    decode_cp_a2a_alltoall,
    decode_cp_a2a_init_workspace,
    decode_cp_a2a_workspace_size,
)
