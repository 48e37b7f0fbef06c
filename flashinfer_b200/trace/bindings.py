"""Which public function carries which template.

``attach()`` runs at package import: with dumping off it only hangs ``.fi_trace`` / ``.__fi_trace_template__`` on the
functions (the objects users call are the undecorated originals); with dumping on - at import through the environment,
or later through :func:`enable` - it swaps the bound names for tracing wrappers, in their home module and wherever
``flashinfer_b200`` re-exports them."""
from __future__ import annotations

import importlib
import sys
from typing import Dict, List, Optional, Tuple

from . import template as _t
from . import templates as T

# (home module, attribute path, template)
BINDINGS: List[Tuple[str, str, "_t.TraceTemplate"]] = [
    ("norm", "rmsnorm", T.rmsnorm_trace),
    ("norm", "fused_add_rmsnorm", T.fused_add_rmsnorm_trace),
    ("norm", "gemma_rmsnorm", T.gemma_rmsnorm_trace),
    ("norm", "gemma_fused_add_rmsnorm", T.gemma_fused_add_rmsnorm_trace),
    ("norm", "layernorm", T.layernorm_trace),
    ("norm", "rmsnorm_quant", T.rmsnorm_quant_trace),
    ("norm", "fused_add_rmsnorm_quant", T.fused_add_rmsnorm_quant_trace),
    ("activation", "silu_and_mul", T.silu_and_mul_trace),
    ("activation", "gelu_and_mul", T.gelu_and_mul_trace),
    ("activation", "gelu_tanh_and_mul", T.gelu_tanh_and_mul_trace),
    ("rope", "apply_rope", T.apply_rope_trace),
    ("rope", "apply_rope_inplace", T.apply_rope_inplace_trace),
    ("rope", "apply_rope_pos_ids", T.apply_rope_pos_ids_trace),
    ("rope", "apply_rope_pos_ids_inplace", T.apply_rope_pos_ids_inplace_trace),
    ("rope", "apply_llama31_rope", T.apply_llama31_rope_trace),
    ("rope", "apply_llama31_rope_inplace", T.apply_llama31_rope_inplace_trace),
    ("rope", "apply_llama31_rope_pos_ids", T.apply_llama31_rope_pos_ids_trace),
    ("rope", "apply_llama31_rope_pos_ids_inplace", T.apply_llama31_rope_pos_ids_inplace_trace),
    ("rope", "apply_rope_with_cos_sin_cache", T.apply_rope_with_cos_sin_cache_trace),
    ("rope", "apply_rope_with_cos_sin_cache_inplace", T.apply_rope_with_cos_sin_cache_inplace_trace),
    ("sampling", "softmax", T.softmax_trace),
    ("sampling", "top_k_renorm_probs", T.top_k_renorm_probs_trace),
    ("sampling", "top_p_renorm_probs", T.top_p_renorm_probs_trace),
    ("sampling", "top_k_mask_logits", T.top_k_mask_logits_trace),
    ("sampling", "sampling_from_probs", T.sampling_from_probs_trace),
    ("sampling", "sampling_from_logits", T.sampling_from_logits_trace),
    ("sampling", "top_k_sampling_from_probs", T.top_k_sampling_from_probs_trace),
    ("sampling", "top_p_sampling_from_probs", T.top_p_sampling_from_probs_trace),
    ("sampling", "min_p_sampling_from_probs", T.min_p_sampling_from_probs_trace),
    ("sampling", "top_k_top_p_sampling_from_probs", T.top_k_top_p_sampling_from_probs_trace),
    ("sampling", "top_k_top_p_sampling_from_logits", T.top_k_top_p_sampling_from_logits_trace),
    ("sampling", "chain_speculative_sampling", T.chain_speculative_sampling_trace),
    ("cascade", "merge_state", T.merge_state_trace),
    ("cascade", "merge_state_in_place", T.merge_state_in_place_trace),
    ("cascade", "merge_states", T.merge_states_trace),
    ("page", "append_paged_kv_cache", T.append_paged_kv_cache_trace),
    ("page", "append_paged_mla_kv_cache", T.append_paged_mla_kv_cache_trace),
    ("page", "get_batch_indices_positions", T.get_batch_indices_positions_trace),
    ("decode", "single_decode_with_kv_cache", T.single_decode_with_kv_cache_trace),
    ("prefill", "single_prefill_with_kv_cache", T.single_prefill_with_kv_cache_trace),
    ("decode", "BatchDecodeWithPagedKVCacheWrapper.run", T.gqa_paged_decode_trace),
    ("prefill", "BatchPrefillWithPagedKVCacheWrapper.run", T.gqa_paged_prefill_trace),
    ("prefill", "BatchPrefillWithRaggedKVCacheWrapper.run", T.gqa_ragged_prefill_trace),
    ("mla", "BatchMLAPagedAttentionWrapper.run", T.mla_paged_trace),
    ("gemm.dense", "mm_bf16", T.mm_bf16_trace),
    ("gemm.dense", "tgv_gemm_sm100", T.tgv_gemm_sm100_trace),
    ("gemm.dense", "bmm_bf16", T.bmm_bf16_trace),
    ("gemm.lowp", "bmm_fp8", T.bmm_fp8_trace),
    ("gemm.lowp", "mm_fp8", T.mm_fp8_trace),
    ("gemm.lowp", "gemm_fp8_nt_groupwise", T.gemm_fp8_nt_groupwise_trace),
    ("gemm.grouped", "SegmentGEMMWrapper.run", T.segment_gemm_trace),
    ("quantization.fp8", "mxfp8_quantize", T.mxfp8_quantize_trace),
    ("quantization.fp4", "fp4_quantize", T.fp4_quantize_trace),
    ("fused_moe.core", "cutlass_fused_moe", T.cutlass_fused_moe_trace),
    ("fused_moe.core", "fused_topk_deepseek", T.fused_topk_deepseek_trace),
    ("topk", "top_k", T.top_k_trace),
    ("concat_ops", "concat_mla_k", T.concat_mla_k_trace),
    ("mamba.selective_state_update", "selective_state_update", T.selective_state_update_trace),
    ("decode", "trtllm_batch_decode_with_kv_cache", T.trtllm_batch_decode_trace),
    ("decode", "cudnn_batch_decode_with_kv_cache", T.cudnn_batch_decode_trace),
    ("gdn", "gated_delta_rule_decode", T.gated_delta_rule_decode_trace),
    ("gdn", "chunk_gated_delta_rule", T.chunk_gated_delta_rule_trace),
    ("rope", "rope_quantize_fp8", T.rope_quantize_fp8_trace),
    ("rope", "mla_rope_quantize_fp8", T.mla_rope_quantize_fp8_trace),
    ("gemm.lowp", "mm_fp4", T.mm_fp4_trace),
    ("gemm.lowp", "mm_mxfp8", T.mm_mxfp8_trace),
    ("gemm.decode_linear", "decode_linear", T.decode_linear_trace),
    ("topk", "topk_clusters_exact", T.topk_clusters_exact_trace),
    ("topk", "top_k_page_table_transform", T.top_k_page_table_transform_trace),
    ("topk", "top_k_ragged_transform", T.top_k_ragged_transform_trace),
    ("triton.sm_constraint_gemm", "gemm_persistent", T.gemm_persistent_trace),
    ("comm.allreduce", "local_argmax", T.local_argmax_trace),
    ("comm.trtllm_alltoall", "moe_local_gather", T.moe_local_gather_trace),
    ("mla._core", "trtllm_batch_decode_with_kv_cache_mla", T.trtllm_batch_decode_mla_trace_dispatch),
    ("sparse", "BlockSparseAttentionWrapper.run", T.block_sparse_attention_trace),
    ("attention._core", "BatchAttention.run", T.batch_attention_trace),
    ("fused_moe.core", "trtllm_bf16_routed_moe", T.trtllm_bf16_routed_moe_trace),
    ("xqa", "xqa", T.xqa_trace),
    ("fused_moe.core", "trtllm_bf16_moe", T.trtllm_bf16_moe_trace),
    ("fused_moe.core", "trtllm_fp8_per_tensor_scale_moe", T.trtllm_fp8_per_tensor_scale_moe_trace),
    ("fused_moe.core", "trtllm_fp8_block_scale_moe", T.trtllm_fp8_block_scale_moe_trace_dispatch),
    ("fused_moe.core", "trtllm_fp8_block_scale_routed_moe", T.trtllm_fp8_block_scale_routed_moe_trace),
    ("fused_moe.core", "trtllm_fp4_block_scale_moe", T.trtllm_fp4_block_scale_moe_trace_dispatch),
    ("fused_moe.core", "trtllm_fp4_block_scale_routed_moe", T.trtllm_fp4_block_scale_routed_moe_trace),
    ("fused_moe.core", "trtllm_mxint4_block_scale_moe", T.trtllm_mxint4_block_scale_moe_trace),
    ("fused_moe.core", "cute_dsl_fused_moe_nvfp4", T.cute_dsl_fused_moe_nvfp4_trace),
    ("fused_moe.core", "CuteDslMoEWrapper.run", T.cute_dsl_moe_wrapper_run_trace),
    ("pod", "PODWithPagedKVCacheWrapper.run", T.pod_with_paged_kv_cache_run_trace),
    ("pod", "BatchPODWithPagedKVCacheWrapper.run", T.batch_pod_with_paged_kv_cache_run_trace),
    ("cascade", "MultiLevelCascadeAttentionWrapper.run", T.multi_level_cascade_run_trace),
    ("sparse", "VariableBlockSparseAttentionWrapper.run", T.variable_block_sparse_attention_run_trace),
    ("prefill", "trtllm_batch_context_with_kv_cache", T.trtllm_batch_context_trace),
    ("prefill", "cudnn_batch_prefill_with_kv_cache", T.cudnn_batch_prefill_trace),
    ("prefill", "trtllm_ragged_attention_deepseek", T.trtllm_ragged_attention_deepseek_trace),
    ("prefill", "fmha_v2_prefill_deepseek", T.fmha_v2_prefill_deepseek_trace),
    ("xqa", "xqa_mla", T.xqa_mla_trace),
    ("decode", "xqa_batch_decode_with_kv_cache", T.xqa_batch_decode_trace),
    ("mla._core", "xqa_batch_decode_with_kv_cache_mla", T.xqa_batch_decode_mla_trace),
    ("norm", "rmsnorm_fp4quant", T.rmsnorm_fp4quant_trace),
    ("norm", "add_rmsnorm_fp4quant", T.add_rmsnorm_fp4quant_trace),
    ("norm", "fused_rmsnorm_silu", T.fused_rmsnorm_silu_trace),
    ("gemm.lowp", "bmm_mxfp8", T.bmm_mxfp8_trace),
    ("gemm.grouped", "batch_deepgemm_fp8_nt_groupwise", T.batch_deepgemm_fp8_nt_groupwise_trace),
    ("gemm.grouped", "grouped_gemm_nt_masked", T.grouped_gemm_nt_masked_trace),
    ("dsv3_ops", "tinygemm_bf16", T.tinygemm_bf16_trace),
    ("quantization.fp4", "nvfp4_quantize", T.nvfp4_quantize_trace),
    ("quantization.fp4", "nvfp4_kv_quantize", T.nvfp4_kv_quantize_trace),
    ("activation", "silu_and_mul_scaled_nvfp4_experts_quantize", T.silu_and_mul_scaled_nvfp4_experts_quantize_trace),
    ("gdn", "gated_delta_rule_mtp", T.gdn_mtp_trace),
    ("mamba.ssd_combined", "SSDCombined.run", T.selective_scan_ssd_prefill_trace),
    ("rope", "rope_quantize_fp8_append_paged_kv_cache", T.rope_quantize_fp8_append_paged_kv_cache_trace),
    ("comm.dcp_alltoall", "decode_cp_a2a_alltoall", T.decode_cp_a2a_alltoall_trace),
]

# one row per concrete template (a dispatch contributes one row per member): what the generic tests iterate over
FLAT_BINDINGS: List[Tuple[str, str, "_t.TraceTemplate"]] = [(m, p, c) for m, p, t in BINDINGS for c in _t.concrete_templates(t)]

_PKG = __name__.rsplit(".", 2)[0]
_originals: Dict[Tuple[str, str], object] = {}


def _resolve(mod_name: str, path: str):
    owner = importlib.import_module(f"{_PKG}.{mod_name}")
    parts = path.split(".")
    for p in parts[:-1]:
        owner = getattr(owner, p)
    return owner, parts[-1]


def _rebind(old, new) -> None:
    """Point every re-export of ``old`` inside the package at ``new``."""
    for name, mod in list(sys.modules.items()):
        if mod is None or not (name == _PKG or name.startswith(_PKG + ".")):
            continue
        for attr, val in list(vars(mod).items()):
            if val is old:
                setattr(mod, attr, new)


def attach() -> int:
    """Bind templates to functions (idempotent).  Returns the number of bound callables."""
    n = 0
    for mod_name, path, tpl in BINDINGS:
        try:
            owner, attr = _resolve(mod_name, path)
            fn = getattr(owner, attr)
        except (ImportError, AttributeError):
            continue
        base = getattr(fn, "__wrapped_untraced__", fn)
        tpl.fi_api = f"{_PKG}.{mod_name}.{path}"
        _originals.setdefault((mod_name, path), base)
        if _t.dump_dir():
            if fn is base:
                wrapped = _t.traced(base, tpl)
                setattr(owner, attr, wrapped)
                if "." not in path:
                    _rebind(base, wrapped)
        else:
            if fn is not base:
                setattr(owner, attr, base)
                if "." not in path:
                    _rebind(fn, base)
            try:
                base.__fi_trace_template__ = tpl
                base.fi_trace = tpl.build_fi_trace_fn()
            except AttributeError:
                pass
        n += 1
    return n


def enable(directory: str) -> None:
    """Start writing one definition file per unique traced call under ``directory``."""
    _t._State.dump_dir = str(directory)
    attach()


def disable() -> None:
    _t._State.dump_dir = None
    attach()


def template_of(fn) -> Optional["_t.TraceTemplate"]:
    fn = getattr(fn, "__func__", fn)
    return getattr(fn, "__fi_trace_template__", None)
