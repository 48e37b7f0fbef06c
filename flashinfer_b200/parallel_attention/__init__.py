"""Context-parallel attention: Ulysses all-to-all + Ring peer-to-peer around the native prefill kernels.

Parity: reference flashinfer/parallel_attention/ (module for module: ``parallel_attention``, ``parallel_config``, ``parallel_wrapper``,
``attention_ops``, ``utils``; same calling conventions - equal-shape padded shards, ``UnevenCPConfig`` / ``VarlenCPConfig``,
``get_parallel_groups`` returning ``(ring_group, ulysses_group)``).  Beyond the reference: causal masking for plain Ulysses-only and
Ring-only runs, fp32 accumulation of the ring partials, GQA shapes with ``fuse_qkv``."""
from . import attention_ops, parallel_attention, parallel_config, parallel_wrapper, utils  # noqa: F401
from .attention_ops import AttentionOpManager  # noqa: F401
from .parallel_attention import ParallelAttention  # noqa: F401
from .parallel_config import UnevenCPConfig, VarlenCPConfig  # noqa: F401
from .utils import (  # noqa: F401
    get_parallel_groups,
    ring_varlen_config,
    split_varlen_input,
    ulysses_varlen_config,
    uneven_cp_config,
)

__all__ = ["ParallelAttention", "UnevenCPConfig", "VarlenCPConfig", "split_varlen_input", "ulysses_varlen_config", "ring_varlen_config",
           "uneven_cp_config", "get_parallel_groups", "AttentionOpManager"]
