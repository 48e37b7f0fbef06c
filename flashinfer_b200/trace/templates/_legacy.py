"""Op templates (one per category; reference flashinfer/trace/templates/*.py)."""
from ..template import Const, Scalar, Tensor, TraceTemplate, Var

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
