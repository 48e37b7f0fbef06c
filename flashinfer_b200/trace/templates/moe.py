"""Mixture-of-experts templates (reference flashinfer/trace/templates/moe.py)."""
import torch

from ..template import Const, Scalar, Tensor, TraceTemplate, Var


def _fused_moe_reference(input, token_selected_experts, token_final_scales, fc1_expert_weights, fc2_expert_weights):
    """fc1 [E, 2I, H] holds [up | gate] rows (the second half is gated through SiLU); fc2 [E, H, I].
    out[t] = sum_j scale[t, j] * fc2[e] @ (silu(gate) * up),  e = token_selected_experts[t, j]."""
    t_, h = input.shape
    inter = fc2_expert_weights.shape[2]
    out = torch.zeros(t_, h, dtype=torch.float32, device=input.device)
    for e in range(fc1_expert_weights.shape[0]):
        tok, slot = torch.nonzero(token_selected_experts == e, as_tuple=True)
        if tok.numel() == 0:
            continue
        hid = input[tok].to(torch.float32) @ fc1_expert_weights[e].to(torch.float32).t()
        act = torch.nn.functional.silu(hid[:, inter:]) * hid[:, :inter]
        y = act @ fc2_expert_weights[e].to(torch.float32).t()
        out.index_add_(0, tok, y * token_final_scales[tok, slot].to(torch.float32)[:, None])
    return out.to(input.dtype)


def _fused_moe_init(*, seq_len=64, num_experts=8, hidden_size=4096, intermediate_size=1024, top_k=2, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = (torch.randn(seq_len, hidden_size, generator=g) * 0.5).to(torch.bfloat16)
    w1 = (torch.randn(num_experts, 2 * intermediate_size, hidden_size, generator=g) / hidden_size ** 0.5).to(torch.bfloat16)
    w2 = (torch.randn(num_experts, hidden_size, intermediate_size, generator=g) / intermediate_size ** 0.5).to(torch.bfloat16)
    scales, ids = torch.topk(torch.softmax(torch.randn(seq_len, num_experts, generator=g), -1), top_k)
    return {"input": x.to(device), "token_selected_experts": ids.int().to(device), "token_final_scales": scales.float().to(device),
            "fc1_expert_weights": w1.to(device), "fc2_expert_weights": w2.to(device), "output_dtype": torch.bfloat16}


def _first(got, expected, kwargs):
    out = got[0][0] if isinstance(got[0], (list, tuple)) else got[0]
    torch.testing.assert_close(out.float(), expected[0].float(), atol=3e-2, rtol=3e-2)


cutlass_fused_moe_trace = TraceTemplate(
    op_type="moe", name_fmt="fused_moe_bf16_e{num_experts}_h{hidden_size}_i{intermediate_size}_topk{top_k}",
    axes=[Var("seq_len"), Const("num_experts", abbrev="e"), Const("hidden_size", abbrev="h"), Const("intermediate_size", abbrev="i"),
          Const("top_k", abbrev="topk")],
    inputs=[Tensor("input", ("seq_len", "hidden_size")), Tensor("token_selected_experts", ("seq_len", "top_k"), "int32"),
            Tensor("token_final_scales", ("seq_len", "top_k"), "float32"),
            Tensor("fc1_expert_weights", ("num_experts", "gate_up_size", "hidden_size")),
            Tensor("fc2_expert_weights", ("num_experts", "hidden_size", "intermediate_size"))],
    outputs=[Tensor("output", ("seq_len", "hidden_size"), dtype_from="input")], reference=_fused_moe_reference, init=_fused_moe_init,
    compare=_first, tags=("moe", "bf16"), constraints=("gate_up_size == 2 * intermediate_size",),
    description="Routed SwiGLU expert FFN with pre-computed routing (bf16 weights)",
    test_sizes={"num_experts": 4, "hidden_size": 64, "intermediate_size": 32, "top_k": 2})


def _fused_topk_deepseek_reference(scores, bias, n_group, topk_group, topk, routed_scaling_factor):
    """DeepSeek-V3 routing: s = sigmoid(scores); pick `topk_group` groups by the sum of each group's two best (s + bias);
    pick `topk` experts by (s + bias) inside those groups; weights = s of the picked experts, normalised, times the factor."""
    s = torch.sigmoid(scores.to(torch.float32))
    sb = s + bias.to(torch.float32)
    t_, e = s.shape
    grp = sb.view(t_, n_group, e // n_group)
    gscore = grp.topk(2, dim=-1).values.sum(-1)
    keep = torch.zeros_like(gscore, dtype=torch.bool).scatter_(1, gscore.topk(topk_group, dim=-1).indices, True)
    masked = torch.where(keep[..., None].expand_as(grp).reshape(t_, e), sb, torch.full_like(sb, float("-inf")))
    idx = masked.topk(topk, dim=-1).indices
    w = s.gather(1, idx)
    return w / w.sum(-1, keepdim=True) * routed_scaling_factor, idx.to(torch.int32)


def _fused_topk_deepseek_init(*, num_tokens=16, num_experts=256, n_group=8, topk_group=4, topk=8, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return {"scores": torch.randn(num_tokens, num_experts, generator=g).to(device), "bias": (torch.randn(num_experts, generator=g) * 0.1).to(device),
            "n_group": n_group, "topk_group": topk_group, "topk": topk, "routed_scaling_factor": 2.5}


def _routing_compare(got, expected, kwargs):
    (w, ids), (w_ref, ids_ref) = got, expected
    order, order_ref = ids.long().sort(-1), ids_ref.long().sort(-1)
    assert torch.equal(order.values, order_ref.values), "selected expert sets differ"
    torch.testing.assert_close(w.float().gather(1, order.indices), w_ref.gather(1, order_ref.indices), atol=1e-5, rtol=1e-4)


fused_topk_deepseek_trace = TraceTemplate(
    op_type="moe", name_fmt="fused_topk_deepseek_e{num_experts}_g{n_group}_tg{topk_group}_k{topk}",
    axes=[Var("num_tokens"), Const("num_experts", abbrev="e"), Const("n_group", abbrev="g"), Const("topk_group", abbrev="tg"), Const("topk", abbrev="k")],
    inputs=[Tensor("scores", ("num_tokens", "num_experts")), Tensor("bias", ("num_experts",)), Scalar("n_group", "int32"),
            Scalar("topk_group", "int32"), Scalar("topk", "int32"), Scalar("routed_scaling_factor")],
    outputs=[Tensor("topk_values", ("num_tokens", "topk"), dtype="float32"), Tensor("topk_indices", ("num_tokens", "topk"), dtype="int32")],
    reference=_fused_topk_deepseek_reference, init=_fused_topk_deepseek_init, compare=_routing_compare, tags=("moe", "routing"),
    description="DeepSeek-V3 no-aux-loss grouped top-k routing",
    test_sizes={"num_experts": 32, "n_group": 4, "topk_group": 2, "topk": 4})
