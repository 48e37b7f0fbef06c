"""API-surface parity: every public name the reference exports at top level (flashinfer/__init__.py) must exist here,
plus the main wrapper classes must expose the reference's public methods.  The list is frozen from the reference tree."""
import pytest

import flashinfer_b200 as fi

REFERENCE_TOP_LEVEL = [
    "ActivationType", "B12xMoEWrapper", "BatchAttention", "BatchAttentionWithAttentionSinkWrapper",
    "BatchDecodeMlaWithPagedKVCacheWrapper", "BatchDecodeWithPagedKVCacheWrapper",
    "BatchDecodeWithSharedPrefixPagedKVCacheWrapper", "BatchMLAPagedAttentionWrapper",
    "BatchPODWithPagedKVCacheWrapper", "BatchPrefillWithPagedKVCacheWrapper", "BatchPrefillWithRaggedKVCacheWrapper",
    "BatchPrefillWithSharedPrefixPagedKVCacheWrapper", "BlockSparseAttentionWrapper",
    "CUDAGraphBatchDecodeWithPagedKVCacheWrapper", "CuteDslMoEWrapper", "MultiLevelCascadeAttentionWrapper",
    "PODWithPagedKVCacheWrapper", "RoutingMethodType", "SegmentGEMMWrapper", "SfLayout", "TopKTieBreak",
    "VariableBlockSparseAttentionWrapper", "add_rmsnorm_fp4quant", "append_paged_kv_cache",
    "append_paged_mla_kv_cache", "apply_llama31_rope", "apply_llama31_rope_inplace", "apply_llama31_rope_pos_ids",
    "apply_llama31_rope_pos_ids_inplace", "apply_rope", "apply_rope_inplace", "apply_rope_pos_ids",
    "apply_rope_pos_ids_inplace", "apply_rope_with_cos_sin_cache", "apply_rope_with_cos_sin_cache_inplace",
    "autotune", "b12x_fused_moe", "block_scale_interleave", "bmm_bf16", "bmm_fp8", "bmm_mxfp8",
    "chain_speculative_sampling", "chunk_gated_delta_rule", "cudnn_batch_decode_with_kv_cache",
    "cute_dsl_fused_moe_nvfp4", "cutlass_fused_moe", "e2m1_and_ufp8sf_scale_to_float", "fast_decode_plan",
    "fi_trace", "fp4_quantize", "fused_add_rmsnorm", "fused_add_rmsnorm_quant", "fused_rmsnorm_silu", "gelu_and_mul",
    "gelu_tanh_and_mul", "gemma_fused_add_rmsnorm", "gemma_rmsnorm", "get_batch_indices_positions",
    "get_fp4_quantization_module", "get_seq_lens", "grouped_mm_bf16", "grouped_mm_fp4", "grouped_mm_fp8",
    "grouped_mm_mxfp8", "jit", "layernorm", "mamba", "merge_state", "merge_state_in_place", "merge_states",
    "min_p_sampling_from_probs", "mm_bf16", "mm_fp4", "mm_fp8", "mm_mxfp8", "mxfp4_dequantize",
    "mxfp4_dequantize_host", "mxfp4_quantize", "mxfp8_dequantize_host", "mxfp8_quantize", "next_positive_power_of_2",
    "nvfp4_batched_quantize", "nvfp4_block_scale_interleave", "nvfp4_kv_dequantize", "nvfp4_kv_quantize",
    "nvfp4_quantize", "nvfp4_quantize_paged_kv_cache", "packbits", "prepare_low_latency_gemm_weights",
    "reorder_rows_for_gated_act_gemm", "rmsnorm", "rmsnorm_fp4quant", "rmsnorm_quant", "sampling_from_logits",
    "sampling_from_probs", "scaled_fp4_grouped_quantize", "segment_packbits", "shuffle_matrix_a",
    "shuffle_matrix_sf_a", "silu_and_mul", "silu_and_mul_scaled_nvfp4_experts_quantize",
    "single_decode_with_kv_cache", "single_prefill_with_kv_cache", "single_prefill_with_kv_cache_return_lse",
    "softmax", "tgv_gemm_sm100", "top_k", "top_k_mask_logits", "top_k_page_table_transform",
    "top_k_ragged_transform", "top_k_renorm_probs", "top_k_sampling_from_probs", "top_k_top_p_sampling_from_logits",
    "top_k_top_p_sampling_from_probs", "top_p_renorm_probs", "top_p_sampling_from_probs", "topk", "trtllm_bf16_moe",
    "trtllm_bf16_routed_moe", "trtllm_fmha_v2_prefill", "trtllm_fp4_block_scale_moe",
    "trtllm_fp4_block_scale_routed_moe", "trtllm_fp8_block_scale_moe", "trtllm_fp8_block_scale_routed_moe",
    "trtllm_fp8_per_tensor_scale_moe", "xqa", "xqa_mla",
]

WRAPPER_METHODS = {
    "BatchDecodeWithPagedKVCacheWrapper": ["plan", "run", "forward", "begin_forward", "end_forward", "reset_workspace_buffer",
                                           "is_cuda_graph_enabled", "use_tensor_cores"],
    "BatchPrefillWithPagedKVCacheWrapper": ["plan", "run", "forward", "begin_forward", "end_forward", "reset_workspace_buffer"],
    "BatchPrefillWithRaggedKVCacheWrapper": ["plan", "run", "forward", "begin_forward", "end_forward", "reset_workspace_buffer"],
    "BatchMLAPagedAttentionWrapper": ["plan", "run"],
    "MultiLevelCascadeAttentionWrapper": ["plan", "run"],
    "PODWithPagedKVCacheWrapper": ["plan", "run", "is_cuda_graph_enabled", "reset_workspace_buffer"],
    "BatchPODWithPagedKVCacheWrapper": ["plan", "run", "is_cuda_graph_enabled"],
    "BlockSparseAttentionWrapper": ["plan", "run", "reset_workspace_buffer"],
    "VariableBlockSparseAttentionWrapper": ["plan", "run", "reset_workspace_buffer"],
    "SegmentGEMMWrapper": ["run", "reset_workspace_buffer"],
    "BatchDecodeMlaWithPagedKVCacheWrapper": ["plan", "run", "is_cuda_graph_enabled", "use_tensor_cores", "reset_workspace_buffer"],
}


def test_reference_top_level_names_exist():
    missing = [n for n in REFERENCE_TOP_LEVEL if not hasattr(fi, n)]
    assert not missing, missing


def test_wrapper_methods_exist():
    missing = [(c, m) for c, ms in WRAPPER_METHODS.items() for m in ms if not hasattr(getattr(fi, c), m)]
    assert not missing, missing


def test_appendix_a_module_level_names():
    """module-level extras and the comm sub-modules."""
    from flashinfer_b200 import comm, decode, diffusion_ops, dsv3_ops, gemm, mla, prefill, rope

    want = {
        decode: ["trtllm_batch_decode_with_kv_cache", "xqa_batch_decode_with_kv_cache"],
        prefill: ["trtllm_batch_context_with_kv_cache", "trtllm_ragged_attention_deepseek", "fmha_varlen", "fmha_varlen_plan",
                  "fmha_v2_prefill_deepseek"],
        mla: ["trtllm_batch_decode_with_kv_cache_mla", "xqa_batch_decode_with_kv_cache_mla"],
        rope: ["rope_quantize_fp8", "mla_rope_quantize_fp8", "rope_quantize_fp8_append_paged_kv_cache"],
        gemm: ["gemm_fp8_nt_groupwise", "gemm_fp8_nt_blockscaled", "group_gemm_fp8_nt_groupwise", "group_gemm_mxfp8_mxfp4_nt_groupwise",
               "group_gemm_nvfp4_nt_groupwise", "group_deepgemm_fp8_nt_groupwise", "batch_deepgemm_fp8_nt_groupwise",
               "fp8_blockscale_gemm_sm90", "trtllm_low_latency_gemm"],
        comm: ["Mapping", "AllReduceFusionOp", "AllReduceFusionPattern", "AllReduceStrategyConfig", "AllReduceStrategyType",
               "QuantizationSFLayout", "trtllm_allreduce_fusion", "trtllm_custom_all_reduce", "trtllm_moe_allreduce_fusion",
               "trtllm_moe_finalize_allreduce_fusion", "trtllm_create_ipc_workspace_for_all_reduce",
               "trtllm_destroy_ipc_workspace_for_all_reduce", "trtllm_create_ipc_workspace_for_all_reduce_fusion",
               "trtllm_destroy_ipc_workspace_for_all_reduce_fusion", "trtllm_lamport_initialize", "trtllm_lamport_initialize_all",
               "compute_fp4_swizzled_layout_sf_size", "allreduce_fusion", "create_allreduce_fusion_workspace",
               "AllReduceFusionWorkspace", "TRTLLMAllReduceFusionWorkspace", "MNNVLAllReduceFusionWorkspace", "MoeAlltoAll",
               "all_gather_matmul", "mnnvl", "nvshmem", "mixed_comm", "trtllm_alltoall", "trtllm_mnnvl_ar", "vllm_all_reduce",
               "vllm_init_custom_ar", "moe_a2a_dispatch", "moe_a2a_combine", "decode_cp_a2a_alltoall"],
        dsv3_ops: ["mm_M1_16_K7168_N128", "mm_M1_16_K7168_N256", "fused_topk_deepseek", "concat_mla_k"],
        diffusion_ops: ["fused_dit_gate_residual_layernorm_gamma_beta"],
    }
    missing = [(m.__name__, n) for m, names in want.items() for n in names if not hasattr(m, n)]
    assert not missing, missing


def test_submodule_entry_points():
    from flashinfer_b200 import comm, fused_moe, gemm, green_ctx, mamba, mla, norm, quantization, sparse, testing, topk

    for mod, names in [
        (comm, ["trtllm_allreduce_fusion", "MoeAlltoAll", "GemmAllReduce", "gemm_reduce_scatter", "AllGatherMatmul", "create_shared_buffer",
                "free_shared_buffer", "pack_strided_memory", "CudaRTLibrary"]),
        (fused_moe, ["trtllm_fp8_block_scale_moe", "moe_forward_fp8_block", "moe_forward_nvfp4", "MoEInputs", "RoutingInputMode",
                     "Fp8QuantizationType", "convert_to_block_layout"]),
        (gemm, ["mm_fp4", "bmm_fp8", "gemm_fp8_nt_groupwise", "group_deepgemm_fp8_nt_groupwise", "batch_deepgemm_fp8_nt_groupwise",
                "linear_gated_silu", "interleave_gate_up", "tinygemm_bf16", "mm_M1_16_K7168_N256"]),
        (green_ctx, ["split_device_green_ctx", "split_device_green_ctx_by_sm_count", "get_cudevice", "get_device_resource",
                     "split_resource", "split_resource_by_sm_count", "create_green_ctx_streams"]),
        (mamba, ["selective_state_update", "SSDCombined"]),
        (mla, ["BatchMLAPagedAttentionWrapper", "MLAHeadDimensions", "MLALayerDimensions", "supported_mla_layer_dimensions"]),
        (norm, ["rmsnorm_fp4quant", "add_rmsnorm_fp4quant", "qk_rmsnorm_cute", "rmsnorm_cute"]),
        (quantization, ["nvfp4_quantize", "nvfp4_quantize_cute_dsl", "get_fp4_quantization_module"]),
        (sparse, ["convert_bsr_mask_layout"]),
        (testing, ["bench_gpu_time", "bench_gpu_time_with_cupti"]),
        (topk, ["top_k", "topk_clusters_exact", "topk_clusters_page_table_transform", "topk_clusters_ragged_transform",
                "can_implement_filtered_topk", "get_fast_topk_clusters"]),
    ]:
        missing = [n for n in names if not hasattr(mod, n)]
        assert not missing, (mod.__name__, missing)


# reference file-level module paths (flashinfer/<path>.py) and one public name each must import from here too
_REFERENCE_MODULE_PATHS = [
    ("gemm.routergemm", "mm_M1_16_K7168_N256"), ("gemm.gemm_base", "mm_fp4"), ("fused_moe.fused_routing_dsv3", "fused_topk_deepseek"),
    ("fused_moe.utils", "get_hybrid_num_tokens_buckets"), ("grouped_mm.core", "grouped_mm_fp8"),
    ("quantization.fp4_quantization", "fp4_quantize"), ("quantization.fp8_quantization", "mxfp8_quantize"),
    ("logits_processor.processors", "TopK"), ("logits_processor.types", "TensorType"), ("logits_processor.compiler", "compile_pipeline"),
    ("parallel_attention.parallel_config", "VarlenCPConfig"), ("parallel_attention.parallel_attention", "ParallelAttention"),
    ("parallel_attention.utils", "get_parallel_groups"), ("cudnn.prefill", "cudnn_batch_prefill_with_kv_cache"),
    ("cudnn.decode", "cudnn_batch_decode_with_kv_cache"), ("version", "__version__"), ("moe_ep", "available_backends"),
    ("comm.trtllm_ar", "trtllm_allreduce_fusion"), ("comm.vllm_ar", "all_reduce"), ("comm.trtllm_moe_alltoall", "MoeAlltoAll"),
    ("fp4_quantization", "fp4_quantize"), ("fp8_quantization", "mxfp8_quantize"), ("gdn_decode", "gated_delta_rule_decode"),
    ("deep_gemm", "m_grouped_fp8_gemm_nt_contiguous"), ("tllm_enums", "ActivationType"),
]


@pytest.mark.parametrize("path,name", _REFERENCE_MODULE_PATHS)
def test_reference_module_paths_import(path, name):
    import importlib

    mod = importlib.import_module(f"flashinfer_b200.{path}")
    assert hasattr(mod, name), f"flashinfer_b200.{path} lacks {name}"


def test_moe_ep_probe_is_truthful():
    from flashinfer_b200 import moe_ep

    assert moe_ep.available_backends() == ["nvlink_a2a"] and not moe_ep.have_nccl_ep() and not moe_ep.have_nixl_ep()
    with pytest.raises(moe_ep.MoEEpNotBuiltError):
        moe_ep.create_fleet(None, 8, 2, 8, 64, backend="nccl_ep")
    with pytest.raises(ValueError):
        moe_ep.create_fleet(None, 8, 2, 8, 64, backend="bogus")
