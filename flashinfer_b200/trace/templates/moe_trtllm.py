"""trtllm-gen style MoE entry points: fused routing + quantised SwiGLU experts (reference flashinfer/trace/templates/moe.py,
the ``trtllm_*_moe`` family; one definition per routing method for the block-scale entry points, chosen per call by a
:class:`TemplateDispatch` on ``routing_method_type`` like the reference's ``*_trace_dispatch`` callables).

Every reference below is routing (written from the routing method's definition, independent of ``fused_moe.route``) followed
by the same fp32 expert FFN over de-quantised weights; the helpers are emitted in front of the reference source in the
definition files so that those stay self-contained."""
import torch

from ..template import Const, Scalar, TemplateDispatch, Tensor, TraceTemplate, Var


# ------------------------------------------------------------------ routing methods (RoutingMethodType values)
def _route_softmax_topk(logits, bias, top_k, n_group, topk_group, scale):
    """Default (0): softmax over all experts, then the top-k probabilities as they are."""
    w, ids = torch.topk(torch.softmax(logits.float(), -1), top_k, -1)
    return ids, w


def _route_topk_softmax(logits, bias, top_k, n_group, topk_group, scale):
    """Renormalize (1): top-k of the raw logits, softmax over the k selected values."""
    v, ids = torch.topk(logits.float(), top_k, -1)
    return ids, torch.softmax(v, -1)


def _route_deepseek_v3(logits, bias, top_k, n_group, topk_group, scale):
    """DeepSeekV3 (2): s = sigmoid(logits); groups ranked by the sum of their two best (s + bias); top-k of (s + bias) inside the
    ``topk_group`` best groups; weights = s of the selected experts / their sum * routed_scaling_factor."""
    s = torch.sigmoid(logits.float())
    sb = s + bias.float()
    t, e = s.shape
    grp = sb.view(t, n_group, e // n_group)
    gscore = grp.topk(2, -1).values.sum(-1)
    keep = torch.zeros_like(gscore, dtype=torch.bool).scatter_(1, gscore.topk(topk_group, -1).indices, True)
    sb = torch.where(keep[..., None].expand_as(grp).reshape(t, e), sb, torch.full_like(sb, float("-inf")))
    ids = sb.topk(top_k, -1).indices
    w = s.gather(1, ids)
    return ids, w / w.sum(-1, keepdim=True) * scale


def _route_llama4(logits, bias, top_k, n_group, topk_group, scale):
    """Llama4 (3): top-k (k = 1 in the model) of the raw logits, weight = sigmoid of the selected logit."""
    v, ids = torch.topk(logits.float(), top_k, -1)
    return ids, torch.sigmoid(v)


def _route_softmax_topk_renorm(logits, bias, top_k, n_group, topk_group, scale):
    """RenormalizeNaive (4): softmax over all experts, top-k, divide by the sum of the k selected probabilities."""
    w, ids = torch.topk(torch.softmax(logits.float(), -1), top_k, -1)
    return ids, w / w.sum(-1, keepdim=True)


def _route_plain_topk(logits, bias, top_k, n_group, topk_group, scale):
    """TopK (5): the k largest scores as they are (the caller normalised them)."""
    w, ids = torch.topk(logits.float(), top_k, -1)
    return ids, w


_ROUTERS = {0: _route_softmax_topk, 1: _route_topk_softmax, 2: _route_deepseek_v3, 3: _route_llama4, 4: _route_softmax_topk_renorm,
            5: _route_plain_topk}


def _swiglu_experts(x, ids, wts, w1, w2, local_expert_offset):
    """x [T, H] fp32; w1 [E_local, 2I, H] = [up | gate] rows; w2 [E_local, H, I]; expert ``e`` of this rank is global expert
    ``e + local_expert_offset`` (tokens routed to other ranks' experts contribute nothing here)."""
    t, h = x.shape
    inter = w2.shape[2]
    out = torch.zeros(t, h, dtype=torch.float32, device=x.device)
    for e in range(w1.shape[0]):
        tok, slot = torch.nonzero(ids == e + local_expert_offset, as_tuple=True)
        if tok.numel() == 0:
            continue
        hid = x[tok] @ w1[e].t()
        act = torch.nn.functional.silu(hid[:, inter:]) * hid[:, :inter]
        out.index_add_(0, tok, (act @ w2[e].t()) * wts[tok, slot][:, None])
    return out


def _dequant_block128(w, scale):
    """e4m3 [..., N, K] with one fp32 scale per 128 x 128 block [..., N/128, K/128]."""
    s = scale.float().repeat_interleave(128, -2).repeat_interleave(128, -1)
    return w.float() * s[..., : w.shape[-2], : w.shape[-1]]


def _dequant_e2m1_block16(packed, sf):
    """packed [..., N, K/2] bytes (two e2m1 values, low nibble first), sf [..., N, K/16] e4m3 bytes in linear layout."""
    mags = [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0]
    lut = torch.tensor(mags + [-m for m in mags], device=packed.device)
    b = packed.view(torch.uint8)
    vals = torch.stack([lut[(b & 0xF).long()], lut[(b >> 4).long()]], -1).flatten(-2)
    s = sf.view(torch.uint8).view(torch.float8_e4m3fn).float().reshape(*vals.shape[:-1], -1)
    return vals * s.repeat_interleave(16, -1)[..., : vals.shape[-1]]


def _unpack_routing_words(topk_ids):
    """int32 word = (expert_id << 16) | bits of the bf16 routing weight."""
    return (topk_ids >> 16).to(torch.int64), (topk_ids & 0xFFFF).to(torch.int16).view(torch.bfloat16).float()


# ------------------------------------------------------------------ shared input builders / compare
def _rand(gen, *shape, scale=1.0):
    return torch.randn(*shape, generator=gen) * scale


def _routing_inputs(gen, method, seq_len, num_experts, top_k, n_group, topk_group):
    logits = _rand(gen, seq_len, num_experts)
    if method == 5:
        logits = torch.softmax(logits, -1)
    bias = _rand(gen, num_experts, scale=0.1).to(torch.bfloat16) if method == 2 else None
    grouped = method == 2
    return {"routing_logits": logits.float() if method == 2 else logits.to(torch.bfloat16), "routing_bias": bias, "num_experts": num_experts,
            "top_k": top_k, "n_group": n_group if grouped else None, "topk_group": topk_group if grouped else None,
            "local_expert_offset": 0, "local_num_experts": num_experts, "routed_scaling_factor": 2.5 if grouped else None,
            "routing_method_type": method}


def _bf16_weights(gen, num_experts, hidden_size, intermediate_size):
    w1 = (_rand(gen, num_experts, 2 * intermediate_size, hidden_size) / hidden_size ** 0.5).to(torch.bfloat16)
    w2 = (_rand(gen, num_experts, hidden_size, intermediate_size) / intermediate_size ** 0.5).to(torch.bfloat16)
    return w1, w2


def _quant_block128(w):
    """fp32 [E, N, K] -> (e4m3, fp32 [E, N/128, K/128]) with amax scaling per block."""
    e, n, k = w.shape
    blocks = w.view(e, n // 128, 128, k // 128, 128)
    s = blocks.abs().amax((2, 4)).clamp_min(1e-6) / 448.0
    q = (blocks / s[:, :, None, :, None]).to(torch.float8_e4m3fn).view(e, n, k)
    return q, s.float()


def _quant_e2m1_block16(w):
    """fp32 [E, N, K] -> (packed bytes [E, N, K/2], e4m3 scale bytes [E, N, K/16]); scale = amax / 6 rounded to e4m3."""
    e, n, k = w.shape
    g = w.view(e, n, k // 16, 16)
    s = (g.abs().amax(-1) / 6.0).clamp_min(2.0 ** -9).to(torch.float8_e4m3fn)
    v = g / s.float()[..., None]
    grid = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0])
    code = (v.abs()[..., None] - grid).abs().argmin(-1) + (v < 0).long() * 8
    code = code.view(e, n, k // 2, 2)
    return (code[..., 0] | (code[..., 1] << 4)).to(torch.uint8), s.view(torch.uint8)


def _to(device, d):
    return {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}


def _close_to_scale(got, expected, kwargs, rel=0.04):
    out = got[0][0] if isinstance(got[0], (list, tuple)) else got[0]
    ref = expected[0].float()
    assert out.shape == ref.shape
    cos = torch.nn.functional.cosine_similarity(out.float().flatten(), ref.flatten(), dim=0)
    assert cos > 0.995, f"cosine similarity {float(cos):.4f}"
    torch.testing.assert_close(out.float(), ref, atol=rel * float(ref.abs().max()), rtol=rel)


def _close_fp4(got, expected, kwargs):
    # the native pipeline re-quantises activations to e2m1 between the two GEMMs: one fp4 step of noise on top of the weights'
    out = got[0][0] if isinstance(got[0], (list, tuple)) else got[0]
    ref = expected[0].float()
    cos = torch.nn.functional.cosine_similarity(out.float().flatten(), ref.flatten(), dim=0)
    assert out.shape == ref.shape and cos > 0.97, f"cosine similarity {float(cos):.4f}"


_MOE_AXES = [Var("seq_len"), Const("num_experts", abbrev="e"), Const("hidden_size", abbrev="h"), Const("intermediate_size", abbrev="i"),
             Const("top_k", abbrev="k")]
_ROUTE_SCALARS = [Scalar("top_k", "int32"), Scalar("n_group", "int32", optional=True), Scalar("topk_group", "int32", optional=True),
                  Scalar("routed_scaling_factor", optional=True), Scalar("local_expert_offset", "int32"),
                  Scalar("routing_method_type", "int32")]
_ROUTE_TENSORS = [Tensor("routing_logits", ("seq_len", "num_experts")), Tensor("routing_bias", ("num_experts",), optional=True)]
_SMALL = {"num_experts": 8, "hidden_size": 128, "intermediate_size": 128, "top_k": 2}
_HELPERS = tuple(_ROUTERS.values()) + (_swiglu_experts,)
_METHOD_NAMES = {0: "default", 1: "renormalize", 2: "ds", 3: "llama4", 4: "renormalize_naive", 5: "topk"}


# ------------------------------------------------------------------ bf16 with fused routing
def _bf16_moe_reference(routing_logits, hidden_states, gemm1_weights, gemm2_weights, top_k, local_expert_offset, routing_method_type,
                        routing_bias=None, n_group=None, topk_group=None, routed_scaling_factor=None):
    ids, wts = _ROUTERS[int(routing_method_type)](routing_logits, routing_bias, top_k, n_group, topk_group, routed_scaling_factor or 1.0)
    out = _swiglu_experts(hidden_states.float(), ids, wts, gemm1_weights.float(), gemm2_weights.float(), local_expert_offset)
    return out.to(hidden_states.dtype)


def _bf16_moe_init(*, seq_len=64, num_experts=8, hidden_size=4096, intermediate_size=1024, top_k=2, device="cuda", seed=0, method=1):
    g = torch.Generator(device="cpu").manual_seed(seed)
    w1, w2 = _bf16_weights(g, num_experts, hidden_size, intermediate_size)
    kw = _routing_inputs(g, method, seq_len, num_experts, top_k, 4, 2)
    kw.update(hidden_states=_rand(g, seq_len, hidden_size, scale=0.5).to(torch.bfloat16), gemm1_weights=w1, gemm2_weights=w2,
              intermediate_size=intermediate_size, use_shuffled_weight=False, weight_layout=0)     # plain K-major weights
    return _to(device, kw)


trtllm_bf16_moe_trace = TraceTemplate(
    op_type="moe", name_fmt="trtllm_bf16_moe_e{num_experts}_h{hidden_size}_i{intermediate_size}_k{top_k}",
    axes=_MOE_AXES + [Var("gemm1_rows")],
    inputs=_ROUTE_TENSORS + [Tensor("hidden_states", ("seq_len", "hidden_size")), Tensor("gemm1_weights", ("num_experts", "gemm1_rows", "hidden_size")),
                             Tensor("gemm2_weights", ("num_experts", "hidden_size", "intermediate_size"))] + _ROUTE_SCALARS,
    outputs=[Tensor("output", ("seq_len", "hidden_size"), dtype_from="hidden_states")], reference=_bf16_moe_reference, init=_bf16_moe_init,
    compare=_close_to_scale, helpers=_HELPERS, tags=("moe", "bf16", "fused-routing"), constraints=("gemm1_rows == 2 * intermediate_size",),
    description="bf16 MoE with routing computed from the router logits inside the op (any RoutingMethodType), SwiGLU experts",
    test_sizes=_SMALL)


# ------------------------------------------------------------------ fp8 per-tensor scales
def _fp8_pt_reference(routing_logits, hidden_states, gemm1_weights, output1_scales_scalar, output1_scales_gate_scalar, gemm2_weights,
                      output2_scales_scalar, top_k, local_expert_offset, routing_method_type, routing_bias=None, n_group=None,
                      topk_group=None, routed_scaling_factor=None):
    """Scale contract: act = silu(gate * s_gate[e]) * (up * s1[e]); out = (act @ W2^T) * s2[e]; the activation scales of the
    e4m3 hidden states / FC2 input are folded into s_gate, s1, s2 by the caller."""
    ids, wts = _ROUTERS[int(routing_method_type)](routing_logits, routing_bias, top_k, n_group, topk_group, routed_scaling_factor or 1.0)
    inter = gemm2_weights.shape[2]
    w1 = gemm1_weights.float().clone()
    w1[:, :inter] *= output1_scales_scalar.float().view(-1, 1, 1)
    w1[:, inter:] *= output1_scales_gate_scalar.float().view(-1, 1, 1)
    w2 = gemm2_weights.float() * output2_scales_scalar.float().view(-1, 1, 1)
    return _swiglu_experts(hidden_states.float(), ids, wts, w1, w2, local_expert_offset).to(torch.bfloat16)


def _fp8_pt_init(*, seq_len=64, num_experts=8, hidden_size=4096, intermediate_size=1024, top_k=2, device="cuda", seed=0, method=1):
    g = torch.Generator(device="cpu").manual_seed(seed)
    w1, w2 = _bf16_weights(g, num_experts, hidden_size, intermediate_size)
    s1 = w1.float().abs().amax((1, 2)) / 448.0
    s2 = w2.float().abs().amax((1, 2)) / 448.0
    x = _rand(g, seq_len, hidden_size, scale=0.5)
    sx = x.abs().max() / 448.0
    kw = _routing_inputs(g, method, seq_len, num_experts, top_k, 4, 2)
    kw.update(hidden_states=(x / sx).to(torch.float8_e4m3fn), gemm1_weights=(w1.float() / s1.view(-1, 1, 1)).to(torch.float8_e4m3fn),
              output1_scales_scalar=(s1 * sx).float(), output1_scales_gate_scalar=(s1 * sx).float(),
              gemm2_weights=(w2.float() / s2.view(-1, 1, 1)).to(torch.float8_e4m3fn), output2_scales_scalar=s2.float(),
              intermediate_size=intermediate_size)
    if kw["routed_scaling_factor"] is None:
        kw["routed_scaling_factor"] = 1.0
    return _to(device, kw)


trtllm_fp8_per_tensor_scale_moe_trace = TraceTemplate(
    op_type="moe", name_fmt="trtllm_fp8_per_tensor_scale_moe_e{num_experts}_h{hidden_size}_i{intermediate_size}_k{top_k}",
    axes=_MOE_AXES + [Var("gemm1_rows")],
    inputs=_ROUTE_TENSORS + [Tensor("hidden_states", ("seq_len", "hidden_size"), "float8_e4m3fn"),
                             Tensor("gemm1_weights", ("num_experts", "gemm1_rows", "hidden_size"), "float8_e4m3fn"),
                             Tensor("output1_scales_scalar", ("num_experts",), "float32"),
                             Tensor("output1_scales_gate_scalar", ("num_experts",), "float32"),
                             Tensor("gemm2_weights", ("num_experts", "hidden_size", "intermediate_size"), "float8_e4m3fn"),
                             Tensor("output2_scales_scalar", ("num_experts",), "float32")] + _ROUTE_SCALARS,
    outputs=[Tensor("output", ("seq_len", "hidden_size"), dtype="bfloat16")], reference=_fp8_pt_reference, init=_fp8_pt_init,
    compare=_close_to_scale, helpers=_HELPERS, tags=("moe", "fp8", "per-tensor", "fused-routing"),
    constraints=("gemm1_rows == 2 * intermediate_size",),
    description="fp8 (e4m3) MoE with one de-quantisation scale per expert and GEMM (up / gate / down), routing inside the op",
    test_sizes=_SMALL)


# ------------------------------------------------------------------ fp8 block scales (DeepSeek layout), one template per routing method
def _fp8_block_reference(routing_logits, hidden_states, hidden_states_scale, gemm1_weights, gemm1_weights_scale, gemm2_weights,
                         gemm2_weights_scale, top_k, local_expert_offset, routing_method_type, routing_bias=None, n_group=None,
                         topk_group=None, routed_scaling_factor=None):
    """hidden_states e4m3 [T, H] with scales [H/128, T] (one per token and 128 channels); weights e4m3 with 128 x 128 block scales."""
    ids, wts = _ROUTERS[int(routing_method_type)](routing_logits, routing_bias, top_k, n_group, topk_group, routed_scaling_factor or 1.0)
    h = hidden_states.shape[1]
    x = hidden_states.float() * hidden_states_scale.float().t().repeat_interleave(128, -1)[:, :h]
    out = _swiglu_experts(x, ids, wts, _dequant_block128(gemm1_weights, gemm1_weights_scale),
                          _dequant_block128(gemm2_weights, gemm2_weights_scale), local_expert_offset)
    return out.to(torch.bfloat16)


def _fp8_block_tensors(g, seq_len, num_experts, hidden_size, intermediate_size):
    w1, w2 = _bf16_weights(g, num_experts, hidden_size, intermediate_size)
    q1, s1 = _quant_block128(w1.float())
    q2, s2 = _quant_block128(w2.float())
    x = _rand(g, seq_len, hidden_size, scale=0.5).view(seq_len, hidden_size // 128, 128)
    sx = x.abs().amax(-1).clamp_min(1e-6) / 448.0
    xq = (x / sx[..., None]).to(torch.float8_e4m3fn).view(seq_len, hidden_size)
    return {"hidden_states": xq, "hidden_states_scale": sx.t().contiguous().float(), "gemm1_weights": q1, "gemm1_weights_scale": s1,
            "gemm2_weights": q2, "gemm2_weights_scale": s2, "intermediate_size": intermediate_size}


def _make_fp8_block_init(method):
    def init(*, seq_len=64, num_experts=8, hidden_size=1024, intermediate_size=512, top_k=2, device="cuda", seed=0):
        g = torch.Generator(device="cpu").manual_seed(seed)
        kw = _routing_inputs(g, method, seq_len, num_experts, top_k, 4, 2)
        kw.update(_fp8_block_tensors(g, seq_len, num_experts, hidden_size, intermediate_size))
        if kw["routed_scaling_factor"] is None:
            kw["routed_scaling_factor"] = 1.0
        return _to(device, kw)

    return init


_FP8_BLOCK_INPUTS = _ROUTE_TENSORS + [
    Tensor("hidden_states", ("seq_len", "hidden_size"), "float8_e4m3fn"), Tensor("hidden_states_scale", ("hidden_blocks", "seq_len"), "float32"),
    Tensor("gemm1_weights", ("num_experts", "gemm1_rows", "hidden_size"), "float8_e4m3fn"),
    Tensor("gemm1_weights_scale", ("num_experts", "gemm1_row_blocks", "hidden_blocks"), "float32"),
    Tensor("gemm2_weights", ("num_experts", "hidden_size", "intermediate_size"), "float8_e4m3fn"),
    Tensor("gemm2_weights_scale", ("num_experts", "hidden_blocks", "inter_blocks"), "float32")]
_FP8_BLOCK_CONSTRAINTS = ("gemm1_rows == 2 * intermediate_size", "hidden_blocks == hidden_size // 128", "inter_blocks == intermediate_size // 128",
                          "gemm1_row_blocks == gemm1_rows // 128")
_BLOCK_AXES = [Var("gemm1_rows"), Var("hidden_blocks"), Var("inter_blocks"), Var("gemm1_row_blocks")]


def _fp8_block_template(method):
    name = _METHOD_NAMES[method]
    return TraceTemplate(
        op_type="moe", name_fmt="trtllm_fp8_block_scale_moe_" + name + "_routing_e{num_experts}_h{hidden_size}_i{intermediate_size}_k{top_k}",
        axes=_MOE_AXES + _BLOCK_AXES, inputs=_FP8_BLOCK_INPUTS + _ROUTE_SCALARS,
        outputs=[Tensor("output", ("seq_len", "hidden_size"), dtype="bfloat16")], reference=_fp8_block_reference,
        init=_make_fp8_block_init(method), compare=_close_to_scale, helpers=_HELPERS + (_dequant_block128,),
        tags=("moe", "fp8", "block-scale", "routing:" + name), constraints=_FP8_BLOCK_CONSTRAINTS,
        description=f"DeepSeek-layout fp8 MoE (1 x 128 activation scales, 128 x 128 weight scales), {_ROUTERS[method].__doc__.split(':')[0]} routing",
        test_sizes=_SMALL)


(trtllm_fp8_block_scale_moe_default_routing_trace, trtllm_fp8_block_scale_moe_renormalize_routing_trace,
 trtllm_fp8_block_scale_moe_ds_routing_trace, trtllm_fp8_block_scale_moe_llama4_routing_trace,
 trtllm_fp8_block_scale_moe_renormalize_naive_routing_trace, trtllm_fp8_block_scale_moe_topk_routing_trace) = (
    _fp8_block_template(m) for m in range(6))

_FP8_BLOCK_BY_METHOD = {0: trtllm_fp8_block_scale_moe_default_routing_trace, 1: trtllm_fp8_block_scale_moe_renormalize_routing_trace,
                        2: trtllm_fp8_block_scale_moe_ds_routing_trace, 3: trtllm_fp8_block_scale_moe_llama4_routing_trace,
                        4: trtllm_fp8_block_scale_moe_renormalize_naive_routing_trace, 5: trtllm_fp8_block_scale_moe_topk_routing_trace}


def _by_routing_method(table):
    def select(bound):
        m = bound.get("routing_method_type", 0)
        return table.get(int(m) if m is not None else 0, table[0])

    return select


trtllm_fp8_block_scale_moe_trace_dispatch = TemplateDispatch(list(_FP8_BLOCK_BY_METHOD.values()), _by_routing_method(_FP8_BLOCK_BY_METHOD))


# ------------------------------------------------------------------ fp8 block scales, packed pre-computed routing
def _fp8_block_routed_reference(topk_ids, hidden_states, hidden_states_scale, gemm1_weights, gemm1_weights_scale, gemm2_weights,
                                gemm2_weights_scale, local_expert_offset):
    ids, wts = _unpack_routing_words(topk_ids)
    h = hidden_states.shape[1]
    x = hidden_states.float() * hidden_states_scale.float().t().repeat_interleave(128, -1)[:, :h]
    out = _swiglu_experts(x, ids, wts, _dequant_block128(gemm1_weights, gemm1_weights_scale),
                          _dequant_block128(gemm2_weights, gemm2_weights_scale), local_expert_offset)
    return out.to(torch.bfloat16)


def _packed_routing(g, seq_len, num_experts, top_k):
    scales, ids = torch.topk(torch.softmax(_rand(g, seq_len, num_experts), -1), top_k)
    return (ids.to(torch.int32) << 16) | (scales.to(torch.bfloat16).view(torch.int16).to(torch.int32) & 0xFFFF)


def _fp8_block_routed_init(*, seq_len=64, num_experts=8, hidden_size=1024, intermediate_size=512, top_k=2, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    kw = {"topk_ids": _packed_routing(g, seq_len, num_experts, top_k), "routing_bias": None, "num_experts": num_experts, "top_k": top_k,
          "n_group": None, "topk_group": None, "local_expert_offset": 0, "local_num_experts": num_experts, "routed_scaling_factor": None}
    kw.update(_fp8_block_tensors(g, seq_len, num_experts, hidden_size, intermediate_size))
    return _to(device, kw)


trtllm_fp8_block_scale_routed_moe_trace = TraceTemplate(
    op_type="moe", name_fmt="trtllm_fp8_block_scale_routed_moe_e{num_experts}_h{hidden_size}_i{intermediate_size}_k{top_k}",
    axes=_MOE_AXES + _BLOCK_AXES,
    inputs=[Tensor("topk_ids", ("seq_len", "top_k"), "int32")] + _FP8_BLOCK_INPUTS[2:] + [Scalar("local_expert_offset", "int32")],
    outputs=[Tensor("output", ("seq_len", "hidden_size"), dtype="bfloat16")], reference=_fp8_block_routed_reference,
    init=_fp8_block_routed_init, compare=_close_to_scale, helpers=(_unpack_routing_words, _swiglu_experts, _dequant_block128),
    tags=("moe", "fp8", "block-scale", "pre-routed"), constraints=_FP8_BLOCK_CONSTRAINTS,
    description="DeepSeek-layout fp8 MoE with packed pre-computed routing ((expert << 16) | bf16 weight)", test_sizes=_SMALL)


# ------------------------------------------------------------------ NVFP4 block scales, one template per routing method
def _fp4_block_reference(routing_logits, hidden_states, gemm1_weights, gemm1_weights_scale, gemm2_weights, gemm2_weights_scale,
                         output1_scale_scalar, output1_scale_gate_scalar, output2_scale_scalar, top_k, local_expert_offset,
                         routing_method_type, routing_bias=None, n_group=None, topk_group=None, routed_scaling_factor=None):
    """e2m1 weights with e4m3 scales per 16 elements; per-expert global scales as in the fp8 per-tensor contract
    (act = silu(gate * s_gate) * (up * s1); out = (act @ W2^T) * s2); bf16 hidden states."""
    ids, wts = _ROUTERS[int(routing_method_type)](routing_logits, routing_bias, top_k, n_group, topk_group, routed_scaling_factor or 1.0)
    w1 = _dequant_e2m1_block16(gemm1_weights, gemm1_weights_scale)
    inter = w1.shape[1] // 2
    w1[:, :inter] *= output1_scale_scalar.float().view(-1, 1, 1)
    w1[:, inter:] *= output1_scale_gate_scalar.float().view(-1, 1, 1)
    w2 = _dequant_e2m1_block16(gemm2_weights, gemm2_weights_scale) * output2_scale_scalar.float().view(-1, 1, 1)
    return _swiglu_experts(hidden_states.float(), ids, wts, w1, w2, local_expert_offset).to(torch.bfloat16)


def _fp4_block_tensors(g, seq_len, num_experts, hidden_size, intermediate_size):
    w1, w2 = _bf16_weights(g, num_experts, hidden_size, intermediate_size)
    g1 = w1.float().abs().amax((1, 2)) / (6.0 * 448.0)
    g2 = w2.float().abs().amax((1, 2)) / (6.0 * 448.0)
    q1, s1 = _quant_e2m1_block16(w1.float() / g1.view(-1, 1, 1))
    q2, s2 = _quant_e2m1_block16(w2.float() / g2.view(-1, 1, 1))
    return {"hidden_states": _rand(g, seq_len, hidden_size, scale=0.5).to(torch.bfloat16), "hidden_states_scale": None, "gemm1_weights": q1,
            "gemm1_weights_scale": s1, "gemm1_bias": None, "gemm1_alpha": None, "gemm1_beta": None, "gemm1_clamp_limit": None,
            "gemm2_weights": q2, "gemm2_weights_scale": s2, "gemm2_bias": None, "output1_scale_scalar": g1.float(),
            "output1_scale_gate_scalar": g1.float(), "output2_scale_scalar": g2.float(), "intermediate_size": intermediate_size}


def _make_fp4_block_init(method):
    def init(*, seq_len=64, num_experts=8, hidden_size=1024, intermediate_size=512, top_k=2, device="cuda", seed=0):
        g = torch.Generator(device="cpu").manual_seed(seed)
        kw = _routing_inputs(g, method, seq_len, num_experts, top_k, 4, 2)
        kw.update(_fp4_block_tensors(g, seq_len, num_experts, hidden_size, intermediate_size))
        if kw["routed_scaling_factor"] is None:
            kw["routed_scaling_factor"] = 1.0
        return _to(device, kw)

    return init


_FP4_WEIGHT_INPUTS = [
    Tensor("hidden_states", ("seq_len", "hidden_size")),
    Tensor("gemm1_weights", ("num_experts", "gemm1_rows", "hidden_half"), "uint8"),
    Tensor("gemm1_weights_scale", ("num_experts", "gemm1_rows", "hidden_groups"), "uint8"),
    Tensor("gemm2_weights", ("num_experts", "hidden_size", "inter_half"), "uint8"),
    Tensor("gemm2_weights_scale", ("num_experts", "hidden_size", "inter_groups"), "uint8"),
    Tensor("output1_scale_scalar", ("num_experts",), "float32"), Tensor("output1_scale_gate_scalar", ("num_experts",), "float32"),
    Tensor("output2_scale_scalar", ("num_experts",), "float32")]
_FP4_CONSTRAINTS = ("gemm1_rows == 2 * intermediate_size", "hidden_half == hidden_size // 2", "hidden_groups == hidden_size // 16",
                    "inter_half == intermediate_size // 2", "inter_groups == intermediate_size // 16")
_FP4_AXES = [Var("gemm1_rows"), Var("hidden_half"), Var("hidden_groups"), Var("inter_half"), Var("inter_groups")]


def _fp4_derive(sizes):
    return {"intermediate_size": sizes["inter_half"] * 2} if "inter_half" in sizes else {}


def _fp4_block_template(method):
    name = _METHOD_NAMES[method]
    return TraceTemplate(
        op_type="moe", name_fmt="trtllm_fp4_block_scale_moe_" + name + "_routing_e{num_experts}_h{hidden_size}_i{intermediate_size}_k{top_k}",
        axes=_MOE_AXES + _FP4_AXES, inputs=_ROUTE_TENSORS + _FP4_WEIGHT_INPUTS + _ROUTE_SCALARS,
        outputs=[Tensor("output", ("seq_len", "hidden_size"), dtype="bfloat16")], reference=_fp4_block_reference,
        init=_make_fp4_block_init(method), compare=_close_fp4, helpers=_HELPERS + (_dequant_e2m1_block16,), derive=_fp4_derive,
        tags=("moe", "nvfp4", "block-scale", "routing:" + name), constraints=_FP4_CONSTRAINTS,
        description=f"NVFP4 MoE (e2m1 weights, e4m3 scales per 16, per-expert global scales), {_ROUTERS[method].__doc__.split(':')[0]} routing",
        test_sizes=_SMALL)


(trtllm_fp4_block_scale_moe_default_routing_trace, trtllm_fp4_block_scale_moe_renormalize_routing_trace,
 trtllm_fp4_block_scale_moe_ds_routing_trace, trtllm_fp4_block_scale_moe_llama4_routing_trace,
 trtllm_fp4_block_scale_moe_renormalize_naive_routing_trace, trtllm_fp4_block_scale_moe_topk_routing_trace) = (
    _fp4_block_template(m) for m in range(6))

_FP4_BLOCK_BY_METHOD = {0: trtllm_fp4_block_scale_moe_default_routing_trace, 1: trtllm_fp4_block_scale_moe_renormalize_routing_trace,
                        2: trtllm_fp4_block_scale_moe_ds_routing_trace, 3: trtllm_fp4_block_scale_moe_llama4_routing_trace,
                        4: trtllm_fp4_block_scale_moe_renormalize_naive_routing_trace, 5: trtllm_fp4_block_scale_moe_topk_routing_trace}
trtllm_fp4_block_scale_moe_trace_dispatch = TemplateDispatch(list(_FP4_BLOCK_BY_METHOD.values()), _by_routing_method(_FP4_BLOCK_BY_METHOD))


def _fp4_block_routed_reference(topk_ids, hidden_states, gemm1_weights, gemm1_weights_scale, gemm2_weights, gemm2_weights_scale,
                                output1_scale_scalar, output1_scale_gate_scalar, output2_scale_scalar, local_expert_offset):
    ids, wts = _unpack_routing_words(topk_ids)
    w1 = _dequant_e2m1_block16(gemm1_weights, gemm1_weights_scale)
    inter = w1.shape[1] // 2
    w1[:, :inter] *= output1_scale_scalar.float().view(-1, 1, 1)
    w1[:, inter:] *= output1_scale_gate_scalar.float().view(-1, 1, 1)
    w2 = _dequant_e2m1_block16(gemm2_weights, gemm2_weights_scale) * output2_scale_scalar.float().view(-1, 1, 1)
    return _swiglu_experts(hidden_states.float(), ids, wts, w1, w2, local_expert_offset).to(torch.bfloat16)


def _fp4_block_routed_init(*, seq_len=64, num_experts=8, hidden_size=1024, intermediate_size=512, top_k=2, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    kw = {"topk_ids": _packed_routing(g, seq_len, num_experts, top_k), "routing_bias": None, "num_experts": num_experts, "top_k": top_k,
          "n_group": None, "topk_group": None, "local_expert_offset": 0, "local_num_experts": num_experts, "routed_scaling_factor": None}
    kw.update(_fp4_block_tensors(g, seq_len, num_experts, hidden_size, intermediate_size))
    return _to(device, kw)


trtllm_fp4_block_scale_routed_moe_trace = TraceTemplate(
    op_type="moe", name_fmt="trtllm_fp4_block_scale_routed_moe_e{num_experts}_h{hidden_size}_i{intermediate_size}_k{top_k}",
    axes=_MOE_AXES + _FP4_AXES,
    inputs=[Tensor("topk_ids", ("seq_len", "top_k"), "int32")] + _FP4_WEIGHT_INPUTS + [Scalar("local_expert_offset", "int32")],
    outputs=[Tensor("output", ("seq_len", "hidden_size"), dtype="bfloat16")], reference=_fp4_block_routed_reference,
    init=_fp4_block_routed_init, compare=_close_fp4, helpers=(_unpack_routing_words, _swiglu_experts, _dequant_e2m1_block16),
    derive=_fp4_derive, tags=("moe", "nvfp4", "block-scale", "pre-routed"), constraints=_FP4_CONSTRAINTS,
    description="NVFP4 MoE with packed pre-computed routing ((expert << 16) | bf16 weight)", test_sizes=_SMALL)


# ------------------------------------------------------------------ MXINT4 weights
def _dequant_int4_block32(packed, scale):
    """packed [..., N, K/2] bytes (two's-complement int4 pairs, low nibble first), scale [..., N, K/32] bf16."""
    b = packed.view(torch.uint8)
    lo, hi = (b & 0xF).to(torch.int16), (b >> 4).to(torch.int16)
    vals = torch.stack([torch.where(lo > 7, lo - 16, lo), torch.where(hi > 7, hi - 16, hi)], -1).flatten(-2).float()
    return vals * scale.float().repeat_interleave(32, -1)[..., : vals.shape[-1]]


def _mxint4_reference(routing_logits, hidden_states, gemm1_weights, gemm1_weights_scale, gemm2_weights, gemm2_weights_scale, top_k,
                      local_expert_offset, routing_method_type, routing_bias=None, n_group=None, topk_group=None, routed_scaling_factor=None):
    ids, wts = _ROUTERS[int(routing_method_type)](routing_logits, routing_bias, top_k, n_group, topk_group, routed_scaling_factor or 1.0)
    out = _swiglu_experts(hidden_states.float(), ids, wts, _dequant_int4_block32(gemm1_weights, gemm1_weights_scale),
                          _dequant_int4_block32(gemm2_weights, gemm2_weights_scale), local_expert_offset)
    return out.to(hidden_states.dtype)


def _quant_int4_block32(w):
    e, n, k = w.shape
    g = w.view(e, n, k // 32, 32)
    s = (g.abs().amax(-1) / 7.0).clamp_min(1e-6).to(torch.bfloat16)
    q = torch.round(g / s.float()[..., None]).clamp(-8, 7).to(torch.int16).view(e, n, k // 2, 2)
    q = torch.where(q < 0, q + 16, q)
    return (q[..., 0] | (q[..., 1] << 4)).to(torch.uint8), s


def _mxint4_init(*, seq_len=64, num_experts=8, hidden_size=1024, intermediate_size=512, top_k=2, device="cuda", seed=0, method=1):
    g = torch.Generator(device="cpu").manual_seed(seed)
    w1, w2 = _bf16_weights(g, num_experts, hidden_size, intermediate_size)
    q1, s1 = _quant_int4_block32(w1.float())
    q2, s2 = _quant_int4_block32(w2.float())
    kw = _routing_inputs(g, method, seq_len, num_experts, top_k, 4, 2)
    kw.update(hidden_states=_rand(g, seq_len, hidden_size, scale=0.5).to(torch.bfloat16), gemm1_weights=q1, gemm1_weights_scale=s1,
              gemm1_alpha=None, gemm1_beta=None, gemm1_clamp_limit=None, gemm2_weights=q2, gemm2_weights_scale=s2,
              intermediate_size=intermediate_size)
    return _to(device, kw)


trtllm_mxint4_block_scale_moe_trace = TraceTemplate(
    op_type="moe", name_fmt="trtllm_mxint4_block_scale_moe_e{num_experts}_h{hidden_size}_i{intermediate_size}_k{top_k}",
    axes=_MOE_AXES + [Var("gemm1_rows"), Var("hidden_half"), Var("hidden_groups"), Var("inter_half"), Var("inter_groups")],
    inputs=_ROUTE_TENSORS + [Tensor("hidden_states", ("seq_len", "hidden_size")),
                             Tensor("gemm1_weights", ("num_experts", "gemm1_rows", "hidden_half"), "uint8"),
                             Tensor("gemm1_weights_scale", ("num_experts", "gemm1_rows", "hidden_groups"), "bfloat16"),
                             Tensor("gemm2_weights", ("num_experts", "hidden_size", "inter_half"), "uint8"),
                             Tensor("gemm2_weights_scale", ("num_experts", "hidden_size", "inter_groups"), "bfloat16")] + _ROUTE_SCALARS,
    outputs=[Tensor("output", ("seq_len", "hidden_size"), dtype_from="hidden_states")], reference=_mxint4_reference, init=_mxint4_init,
    compare=_close_to_scale, helpers=_HELPERS + (_dequant_int4_block32,), derive=_fp4_derive, tags=("moe", "mxint4", "fused-routing"),
    constraints=("gemm1_rows == 2 * intermediate_size", "hidden_half == hidden_size // 2", "hidden_groups == hidden_size // 32",
                 "inter_half == intermediate_size // 2", "inter_groups == intermediate_size // 32"),
    description="MoE with int4 weights and one bf16 scale per 32 elements (expanded to bf16 once at load time on Blackwell)",
    test_sizes=_SMALL)


# ------------------------------------------------------------------ NVFP4 with pre-computed routing (cute-dsl style entry point + wrapper)
def _cute_nvfp4_reference(x, token_selected_experts, token_final_scales, w1_weight, w1_weight_sf, w1_alpha, w2_weight, w2_weight_sf, w2_alpha):
    w1 = _dequant_e2m1_block16(w1_weight, w1_weight_sf) * w1_alpha.float().view(-1, 1, 1)
    w2 = _dequant_e2m1_block16(w2_weight, w2_weight_sf) * w2_alpha.float().view(-1, 1, 1)
    return _swiglu_experts(x.float(), token_selected_experts.long(), token_final_scales.float(), w1, w2, 0).to(torch.bfloat16)


def _cute_nvfp4_kwargs(seq_len, num_experts, hidden_size, intermediate_size, top_k, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    t = _fp4_block_tensors(g, seq_len, num_experts, hidden_size, intermediate_size)
    scales, ids = torch.topk(torch.softmax(_rand(g, seq_len, num_experts), -1), top_k)
    return {"x": t["hidden_states"], "x_sf": None, "token_selected_experts": ids.int(), "token_final_scales": scales.float(),
            "w1_weight": t["gemm1_weights"], "w1_weight_sf": t["gemm1_weights_scale"], "w1_alpha": t["output1_scale_scalar"],
            "fc2_input_scale": None, "w2_weight": t["gemm2_weights"], "w2_weight_sf": t["gemm2_weights_scale"],
            "w2_alpha": t["output2_scale_scalar"]}


def _cute_nvfp4_init(*, seq_len=64, num_experts=8, hidden_size=1024, intermediate_size=512, top_k=2, device="cuda", seed=0):
    kw = _cute_nvfp4_kwargs(seq_len, num_experts, hidden_size, intermediate_size, top_k, seed)
    kw.update(num_experts=num_experts, top_k=top_k)
    return _to(device, kw)


def _cute_wrapper_init(*, seq_len=64, num_experts=8, hidden_size=1024, intermediate_size=512, top_k=2, device="cuda", seed=0):
    from ...fused_moe.core import CuteDslMoEWrapper

    kw = _to(device, _cute_nvfp4_kwargs(seq_len, num_experts, hidden_size, intermediate_size, top_k, seed))
    kw["self"] = CuteDslMoEWrapper(num_experts, top_k, hidden_size, intermediate_size)
    return kw


_CUTE_INPUTS = [Tensor("x", ("seq_len", "hidden_size")), Tensor("token_selected_experts", ("seq_len", "top_k"), "int32"),
                Tensor("token_final_scales", ("seq_len", "top_k"), "float32"),
                Tensor("w1_weight", ("num_experts", "gemm1_rows", "hidden_half"), "uint8"),
                Tensor("w1_weight_sf", ("num_experts", "gemm1_rows", "hidden_groups"), "uint8"), Tensor("w1_alpha", ("num_experts",), "float32"),
                Tensor("w2_weight", ("num_experts", "hidden_size", "inter_half"), "uint8"),
                Tensor("w2_weight_sf", ("num_experts", "hidden_size", "inter_groups"), "uint8"), Tensor("w2_alpha", ("num_experts",), "float32")]

cute_dsl_fused_moe_nvfp4_trace = TraceTemplate(
    op_type="moe", name_fmt="cute_dsl_fused_moe_nvfp4_e{num_experts}_h{hidden_size}_i{intermediate_size}_k{top_k}",
    axes=_MOE_AXES + _FP4_AXES, inputs=_CUTE_INPUTS, outputs=[Tensor("output", ("seq_len", "hidden_size"), dtype="bfloat16")],
    reference=_cute_nvfp4_reference, init=_cute_nvfp4_init, compare=_close_fp4, helpers=(_swiglu_experts, _dequant_e2m1_block16),
    derive=_fp4_derive, tags=("moe", "nvfp4", "pre-routed"), constraints=_FP4_CONSTRAINTS,
    description="NVFP4 MoE with pre-computed routing (ids + weights) on the block-scaled grouped tcgen05 GEMMs", test_sizes=_SMALL)

cute_dsl_moe_wrapper_run_trace = TraceTemplate(
    op_type="moe", name_fmt="cute_dsl_moe_wrapper_e{num_experts}_h{hidden_size}_i{intermediate_size}_k{top_k}",
    axes=_MOE_AXES + _FP4_AXES, inputs=_CUTE_INPUTS, outputs=[Tensor("output", ("seq_len", "hidden_size"), dtype="bfloat16")],
    reference=_cute_nvfp4_reference, init=_cute_wrapper_init, compare=_close_fp4, helpers=(_swiglu_experts, _dequant_e2m1_block16),
    derive=_fp4_derive, tags=("moe", "nvfp4", "pre-routed", "wrapper"), constraints=_FP4_CONSTRAINTS,
    description="Stateful wrapper form of the NVFP4 pre-routed MoE (expert-parallel placement bound at construction)", test_sizes=_SMALL)

__all__ = [n for n in dir() if n.endswith("_trace") or n.endswith("_trace_dispatch")]
