"""Op templates (one per category; reference flashinfer/trace/templates/*.py)."""
from ..template import Const, Scalar, Tensor, TraceTemplate, Var

gemm_bf16_trace = TraceTemplate(
    op_type="gemm", name_fmt="gemm_bf16_n{N}_k{K}", axes=[Var("M"), Const("N"), Const("K")],
    inputs=[Tensor("a", ("M", "K")), Tensor("b", ("K", "N"))], outputs=[Tensor("out", ("M", "N"))], tags=("gemm", "bf16"),
    description="C = A @ B with a column-major B ([N, K] weight passed as .T)")

gqa_paged_decode_trace = TraceTemplate(
    op_type="gqa_paged", name_fmt="gqa_paged_decode_h{num_qo_heads}_kv{num_kv_heads}_d{head_dim}_ps{page_size}",
    axes=[Var("batch_size"), Const("num_qo_heads"), Const("num_kv_heads"), Const("head_dim"), Const("page_size"),
          Var("num_pages"), Var("len_indptr"), Var("num_kv_indices")],
    inputs=[Tensor("q", ("batch_size", "num_qo_heads", "head_dim")),
            Tensor("k_cache", ("num_pages", "page_size", "num_kv_heads", "head_dim")),
            Tensor("v_cache", ("num_pages", "page_size", "num_kv_heads", "head_dim")),
            Tensor("kv_indptr", ("len_indptr",), "int32"), Tensor("kv_indices", ("num_kv_indices",), "int32"),
            Tensor("kv_last_page_len", ("batch_size",), "int32"), Scalar("sm_scale")],
    outputs=[Tensor("output", ("batch_size", "num_qo_heads", "head_dim")), Tensor("lse", ("batch_size", "num_qo_heads"), "float32")],
    tags=("attention", "decode"), description="Batched GQA decode over a paged KV cache")

gqa_ragged_prefill_trace = TraceTemplate(
    op_type="gqa_ragged", name_fmt="gqa_ragged_prefill_causal_h{num_qo_heads}_kv{num_kv_heads}_d{head_dim}",
    axes=[Var("total_q"), Var("total_kv"), Const("num_qo_heads"), Const("num_kv_heads"), Const("head_dim"), Var("len_indptr")],
    inputs=[Tensor("q", ("total_q", "num_qo_heads", "head_dim")), Tensor("k", ("total_kv", "num_kv_heads", "head_dim")),
            Tensor("v", ("total_kv", "num_kv_heads", "head_dim")), Tensor("qo_indptr", ("len_indptr",), "int32"),
            Tensor("kv_indptr", ("len_indptr",), "int32"), Scalar("sm_scale")],
    outputs=[Tensor("output", ("total_q", "num_qo_heads", "head_dim")), Tensor("lse", ("total_q", "num_qo_heads"), "float32")],
    tags=("attention", "prefill"), description="Batched causal GQA prefill over ragged Q/K/V")

mla_paged_decode_trace = TraceTemplate(
    op_type="mla_paged", name_fmt="mla_paged_decode_h{num_heads}_ckv{head_dim_ckv}_kpe{head_dim_kpe}_ps{page_size}",
    axes=[Var("batch_size"), Const("num_heads"), Const("head_dim_ckv"), Const("head_dim_kpe"), Const("page_size"), Var("num_pages")],
    inputs=[Tensor("q_nope", ("batch_size", "num_heads", "head_dim_ckv")), Tensor("q_pe", ("batch_size", "num_heads", "head_dim_kpe")),
            Tensor("ckv_cache", ("num_pages", "page_size", "head_dim_ckv")), Tensor("kpe_cache", ("num_pages", "page_size", "head_dim_kpe"))],
    outputs=[Tensor("output", ("batch_size", "num_heads", "head_dim_ckv"))], tags=("attention", "mla"),
    description="Multi-head latent attention decode (matrix-absorbed)")

sampling_trace = TraceTemplate(
    op_type="sampling", name_fmt="top_k_top_p_sampling_v{vocab_size}", axes=[Var("batch_size"), Const("vocab_size")],
    inputs=[Tensor("probs", ("batch_size", "vocab_size"), "float32")], outputs=[Tensor("samples", ("batch_size",), "int32")],
    tags=("sampling",), description="Sorting-free rejection sampling with top-k / top-p filtering")

moe_trace = TraceTemplate(
    op_type="moe", name_fmt="moe_bf16_e{num_experts}_h{hidden_size}_i{intermediate_size}_topk{top_k}",
    axes=[Var("seq_len"), Const("num_experts"), Const("hidden_size"), Const("intermediate_size"), Const("top_k")],
    inputs=[Tensor("routing_logits", ("seq_len", "num_experts")), Tensor("hidden_states", ("seq_len", "hidden_size"))],
    outputs=[Tensor("output", ("seq_len", "hidden_size"))], tags=("moe",), description="Routed mixture-of-experts FFN")

allreduce_fusion_trace = TraceTemplate(
    op_type="comm", name_fmt="allreduce_add_rmsnorm_h{hidden_size}", axes=[Var("num_tokens"), Const("hidden_size")],
    inputs=[Tensor("input", ("num_tokens", "hidden_size")), Tensor("residual", ("num_tokens", "hidden_size")),
            Tensor("weight", ("hidden_size",))],
    outputs=[Tensor("norm_out", ("num_tokens", "hidden_size")), Tensor("residual_out", ("num_tokens", "hidden_size"))],
    tags=("comm", "tp"), description="Tensor-parallel all-reduce fused with residual add and RMSNorm")
