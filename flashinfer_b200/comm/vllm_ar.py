"""Module path of the reference (flashinfer/comm/vllm_ar.py): vLLM-style custom all-reduce handle API (implementation: compat.py)."""
from .compat import vllm_all_reduce as all_reduce  # noqa: F401
from .compat import vllm_dispose as dispose  # noqa: F401
from .compat import vllm_get_graph_buffer_ipc_meta as get_graph_buffer_ipc_meta  # noqa: F401
from .compat import vllm_init_custom_ar as init_custom_ar  # noqa: F401
from .compat import vllm_meta_size as meta_size  # noqa: F401
from .compat import vllm_register_buffer as register_buffer  # noqa: F401
from .compat import vllm_register_graph_buffers as register_graph_buffers  # noqa: F401
