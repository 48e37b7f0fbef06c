"""Communication kernels over NVLink 5 / NVSwitch (symmetric heap + in-kernel collectives)."""
from .allreduce import TPCommunicator  # noqa: F401
from .symm import SymmetricHeap  # noqa: F401
from .mapping import Mapping  # noqa: F401,E402
from .moe_alltoall import (  # noqa: F401,E402
    MoeAlltoAll,
    moe_a2a_combine,
    moe_a2a_dispatch,
    moe_a2a_get_workspace_size_per_rank,
    moe_a2a_initialize,
    moe_a2a_sanitize_expert_ids,
    moe_a2a_wrap_payload_tensor_in_workspace,
)
