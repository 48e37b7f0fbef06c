"""Reference import path flashinfer/cute_dsl/attention: variant objects and the CuTe-DSL wrapper classes.  The DSL building blocks
of the reference (warp schedules, TMEM layouts, pipeline topologies, tile schedulers) describe ITS kernel; the equivalents of this
library are the C++ structures inside csrc/attention/{prefill,mla}_sm100.cu and the planners in csrc/runtime/planner.cpp."""
from .variant import (  # noqa: F401
    ALiBiAttention,
    AttentionVariant,
    AttentionWithSink,
    RPEAttention,
    SigmoidAttention,
    SigmoidTanhAttention,
    SoftCappingAttention,
    StandardAttention,
    tanh_approx,
)
from .wrappers import BatchMLADecodeCuteDSLWrapper, BatchPrefillCuteDSLWrapper, cute_dsl_mla_decode  # noqa: F401


def mla_get_split_kv(batch_size: int, q_len: int, max_seq_len: int, num_sms: int) -> int:
    """How many KV splits the MLA decode planner would use for a uniform batch (reference scheduler/mla_persistent.py
    ``mla_get_split_kv``): enough splits to occupy the SM pairs, never finer than one 64-token page per split."""
    work = max(1, batch_size * q_len)
    pairs = max(1, num_sms // 2)
    return max(1, min(pairs // work if work < pairs else 1, -(-max_seq_len // 64)))


def mla_get_workspace_size(batch_size: int, q_len: int, num_heads: int, kv_lora_rank: int, split_kv: int) -> int:
    """Bytes of split-KV scratch: fp32 partial outputs + one log-sum-exp per (request, token, head, split); 0 without splitting."""
    return 0 if split_kv <= 1 else batch_size * q_len * num_heads * split_kv * (kv_lora_rank + 1) * 4
