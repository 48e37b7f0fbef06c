"""Communication templates (reference flashinfer/trace/templates/comm.py): schema only - the reference needs a process group."""
from ..template import Const, Scalar, Tensor, TraceTemplate, Var

allreduce_fusion_trace = TraceTemplate(
    op_type="comm", name_fmt="allreduce_add_rmsnorm_h{hidden_size}", axes=[Var("num_tokens"), Const("hidden_size")],
    inputs=[Tensor("input", ("num_tokens", "hidden_size")), Tensor("residual", ("num_tokens", "hidden_size")),
            Tensor("weight", ("hidden_size",))],
    outputs=[Tensor("norm_out", ("num_tokens", "hidden_size")), Tensor("residual_out", ("num_tokens", "hidden_size"))],
    tags=("comm", "tp"), description="Tensor-parallel all-reduce fused with residual add and RMSNorm")
