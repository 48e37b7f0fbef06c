"""Communication kernels over NVLink 5 / NVSwitch (symmetric heap + in-kernel collectives)."""
from .allreduce import TPCommunicator  # noqa: F401
from .symm import SymmetricHeap  # noqa: F401
from .mapping import Mapping  # noqa: F401,E402
from .trtllm_moe_alltoall import (  # noqa: F401,E402
    MoeAlltoAll,
    moe_a2a_combine,
    moe_a2a_dispatch,
    moe_a2a_get_workspace_size_per_rank,
    moe_a2a_initialize,
    moe_a2a_sanitize_expert_ids,
    moe_a2a_wrap_payload_tensor_in_workspace,
)
from .workspace_base import AllReduceFusionWorkspace  # noqa: F401,E402
from .trtllm_ar import (  # noqa: F401,E402
    AllReduceFusionOp,
    AllReduceFusionPattern,
    AllReduceStrategyConfig,
    AllReduceStrategyType,
    QuantizationSFLayout,
    TRTLLMAllReduceFusionWorkspace,
    allreduce_fusion,
    compute_fp4_swizzled_layout_sf_size,
    trtllm_allreduce_fusion,
    trtllm_create_ipc_workspace_for_all_reduce,
    trtllm_create_ipc_workspace_for_all_reduce_fusion,
    trtllm_custom_all_reduce,
    trtllm_destroy_ipc_workspace_for_all_reduce,
    trtllm_destroy_ipc_workspace_for_all_reduce_fusion,
    trtllm_lamport_initialize,
    trtllm_lamport_initialize_all,
    trtllm_moe_allreduce_fusion,
    trtllm_moe_finalize_allreduce_fusion,
)
from .allreduce import create_allreduce_fusion_workspace  # noqa: F401,E402
from .trtllm_mnnvl_ar import MNNVLAllReduceFusionWorkspace  # noqa: F401,E402
from .vllm_ar import (  # noqa: F401,E402
    vllm_all_reduce,
    vllm_dispose,
    vllm_get_graph_buffer_ipc_meta,
    vllm_init_custom_ar,
    vllm_meta_size,
    vllm_register_buffer,
    vllm_register_graph_buffers,
)
from .collectives import NVLSCollectives  # noqa: F401,E402
from .mixed_comm import MixedCommHandler, MixedCommMode, MixedCommOp, run_mixed_comm  # noqa: F401,E402
from .dcp_alltoall import (  # noqa: F401,E402
    decode_cp_a2a_allocate_mnnvl_workspace,
    decode_cp_a2a_alltoall,
    decode_cp_a2a_init_workspace,
    decode_cp_a2a_workspace_size,
)
from .gemm_allreduce import GemmAllReduce, gemm_allreduce, gemm_reduce_scatter  # noqa: F401,E402
from .all_gather_matmul import AllGatherMatmul, all_gather_matmul  # noqa: F401,E402
from .cuda_ipc import CudaRTLibrary, create_shared_buffer, free_shared_buffer  # noqa: F401,E402
from .dlpack_utils import pack_strided_memory  # noqa: F401,E402
from . import mixed_comm, mnnvl, nvshmem, trtllm_alltoall, trtllm_mnnvl_ar  # noqa: F401,E402
from .trtllm_mnnvl_ar import (  # noqa: F401,E402
    MNNVLAllreduceFusionStrategy,
    trtllm_mnnvl_allreduce,
    trtllm_mnnvl_fused_allreduce_add_rmsnorm,
)
from . import (  # noqa: F401,E402
    cuda_ipc,
    dcp_alltoall,
    dlpack_utils,
    nvshmem_allreduce,
    torch_symmetric_memory,
    trtllm_ar,
    trtllm_moe_alltoall,
    vllm_ar,
    workspace_base,
)
