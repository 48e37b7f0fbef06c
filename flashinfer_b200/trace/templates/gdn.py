"""Gated delta rule (Qwen3-Next / GDN linear attention) templates (reference flashinfer/trace/templates/gdn.py).

Per value head with state S [K, V]:   S <- g S;   S <- S + k (beta (v - k^T S))^T;   o = (scale q)^T S,
g = exp(-exp(A_log) softplus(a + dt_bias)) and beta = sigmoid(b) on the decode path, given directly on the prefill path."""
import torch

from ..template import Const, Scalar, Tensor, TraceTemplate, Var

_AXES = [Const("num_q_heads", abbrev="h"), Const("num_v_heads", abbrev="hv"), Const("head_dim_k", abbrev="k"), Const("head_dim_v", abbrev="v")]
_SIZES = {"num_q_heads": 2, "num_v_heads": 4, "head_dim_k": 16, "head_dim_v": 8}


def _gdn_decode_reference(q, k, v, state, A_log, a, dt_bias, b, scale=None, use_qk_l2norm=True):
    """q, k [B, 1, H, K]; v [B, 1, HV, V]; state [B, HV, K, V] fp32; A_log, dt_bias [HV]; a, b [B, 1, HV]."""
    bsz, _, h, kd = q.shape
    hv = v.shape[2]
    rep = hv // h
    sc = scale if scale is not None else kd ** -0.5
    g = torch.exp(-torch.exp(A_log.to(torch.float32)) * torch.nn.functional.softplus(a.to(torch.float32) + dt_bias.to(torch.float32)))
    beta = torch.sigmoid(b.to(torch.float32))
    out = torch.zeros(bsz, 1, hv, v.shape[3], dtype=torch.float32, device=q.device)
    new_state = state.to(torch.float32).clone()
    for i in range(bsz):
        qt = q[i, 0].to(torch.float32).repeat_interleave(rep, 0)
        kt = k[i, 0].to(torch.float32).repeat_interleave(rep, 0)
        if use_qk_l2norm:
            qt = qt * torch.rsqrt((qt * qt).sum(-1, keepdim=True) + 1e-6)
            kt = kt * torch.rsqrt((kt * kt).sum(-1, keepdim=True) + 1e-6)
        s = new_state[i] * g[i, 0][:, None, None]
        delta = (v[i, 0].to(torch.float32) - torch.einsum("hk,hkv->hv", kt, s)) * beta[i, 0][:, None]
        s = s + kt[:, :, None] * delta[:, None, :]
        out[i, 0] = torch.einsum("hk,hkv->hv", qt * sc, s)
        new_state[i] = s
    return out.to(q.dtype), new_state


def _gdn_decode_init(*, batch_size=8, num_q_heads=16, num_v_heads=32, head_dim_k=128, head_dim_v=128, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    return {"q": r(batch_size, 1, num_q_heads, head_dim_k).to(torch.bfloat16).to(device), "k": r(batch_size, 1, num_q_heads, head_dim_k).to(torch.bfloat16).to(device),
            "v": r(batch_size, 1, num_v_heads, head_dim_v).to(torch.bfloat16).to(device),
            "state": (r(batch_size, num_v_heads, head_dim_k, head_dim_v) * 0.1).to(device), "A_log": (r(num_v_heads) * 0.5).to(device),
            "a": r(batch_size, 1, num_v_heads).to(torch.bfloat16).to(device), "dt_bias": (r(num_v_heads) * 0.1).to(device),
            "b": r(batch_size, 1, num_v_heads).to(torch.bfloat16).to(device)}


gated_delta_rule_decode_trace = TraceTemplate(
    op_type="gdn", name_fmt="gdn_decode_h{num_q_heads}_hv{num_v_heads}_k{head_dim_k}_v{head_dim_v}", axes=[Var("batch_size")] + _AXES,
    inputs=[Tensor("q", ("batch_size", "one", "num_q_heads", "head_dim_k")), Tensor("k", ("batch_size", "one", "num_q_heads", "head_dim_k")),
            Tensor("v", ("batch_size", "one", "num_v_heads", "head_dim_v")), Tensor("state", ("batch_size", "num_v_heads", "head_dim_k", "head_dim_v"), "float32"),
            Tensor("A_log", ("num_v_heads",)), Tensor("a", ("batch_size", "one", "num_v_heads")), Tensor("dt_bias", ("num_v_heads",)),
            Tensor("b", ("batch_size", "one", "num_v_heads")), Scalar("scale", optional=True), Scalar("use_qk_l2norm", "bool", optional=True)],
    outputs=[Tensor("output", ("batch_size", "one", "num_v_heads", "head_dim_v"), dtype_from="q"),
             Tensor("state_out", ("batch_size", "num_v_heads", "head_dim_k", "head_dim_v"), dtype="float32", param="state")],
    reference=_gdn_decode_reference, init=_gdn_decode_init, tags=("gdn", "decode", "inplace"), constraints=("one == 1",),
    description="Gated delta rule, one token per sequence, K-major fp32 state updated in place", tolerance="bf16", test_sizes=_SIZES)


def _gdn_prefill_reference(q, k, v, g, beta, cu_seqlens, scale=None, initial_state=None, use_qk_l2norm_in_kernel=False):
    """q, k [total, H, K]; v [total, HV, V]; g (multiplicative gate), beta [total, HV]; packed sequences in cu_seqlens.
    States are K-last: initial_state and the returned final state [num_seqs, HV, V, K].  Returns (output [total, HV, V], final state)."""
    total, h, kd = q.shape
    hv, vd = v.shape[1], v.shape[2]
    rep = hv // h
    sc = scale if scale is not None else kd ** -0.5
    n = cu_seqlens.numel() - 1
    out = torch.zeros(total, hv, vd, dtype=torch.float32, device=q.device)
    states = torch.zeros(n, hv, kd, vd, dtype=torch.float32, device=q.device) if initial_state is None else initial_state.to(torch.float32).transpose(-1, -2).clone()
    for i in range(n):
        s = states[i]
        for t in range(int(cu_seqlens[i]), int(cu_seqlens[i + 1])):
            qt = q[t].to(torch.float32).repeat_interleave(rep, 0)
            kt = k[t].to(torch.float32).repeat_interleave(rep, 0)
            if use_qk_l2norm_in_kernel:
                qt = qt * torch.rsqrt((qt * qt).sum(-1, keepdim=True) + 1e-6)
                kt = kt * torch.rsqrt((kt * kt).sum(-1, keepdim=True) + 1e-6)
            s = s * g[t].to(torch.float32)[:, None, None]
            delta = (v[t].to(torch.float32) - torch.einsum("hk,hkv->hv", kt, s)) * beta[t].to(torch.float32)[:, None]
            s = s + kt[:, :, None] * delta[:, None, :]
            out[t] = torch.einsum("hk,hkv->hv", qt * sc, s)
        states[i] = s
    return out.to(q.dtype), states.transpose(-1, -2).contiguous()


def _gdn_prefill_init(*, num_seqs=4, seq_len=None, num_q_heads=16, num_v_heads=32, head_dim_k=128, head_dim_v=128, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    lens = [int(x) for x in torch.randint(3, 90, (num_seqs,), generator=g)] if seq_len is None else [int(seq_len)] * num_seqs
    total = sum(lens)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    kk = r(total, num_q_heads, head_dim_k)
    kk = kk / kk.norm(dim=-1, keepdim=True)                      # unit keys keep the delta rule contractive
    return {"q": r(total, num_q_heads, head_dim_k).to(torch.bfloat16).to(device), "k": kk.to(torch.bfloat16).to(device),
            "v": r(total, num_v_heads, head_dim_v).to(torch.bfloat16).to(device),
            "g": torch.exp(-torch.rand(total, num_v_heads, generator=g) * 0.2).to(device), "beta": torch.rand(total, num_v_heads, generator=g).to(device),
            "cu_seqlens": torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=device), "output_final_state": True}


chunk_gated_delta_rule_trace = TraceTemplate(
    op_type="gdn", name_fmt="gdn_prefill_h{num_q_heads}_hv{num_v_heads}_k{head_dim_k}_v{head_dim_v}",
    axes=[Var("total_tokens"), Var("num_seqs"), Var("len_cu")] + _AXES,
    inputs=[Tensor("q", ("total_tokens", "num_q_heads", "head_dim_k")), Tensor("k", ("total_tokens", "num_q_heads", "head_dim_k")),
            Tensor("v", ("total_tokens", "num_v_heads", "head_dim_v")), Tensor("g", ("total_tokens", "num_v_heads"), "float32"),
            Tensor("beta", ("total_tokens", "num_v_heads"), "float32"), Tensor("cu_seqlens", ("len_cu",), "int32"), Scalar("scale", optional=True),
            Tensor("initial_state", ("num_seqs", "num_v_heads", "head_dim_v", "head_dim_k"), "float32", optional=True),
            Scalar("use_qk_l2norm_in_kernel", "bool", optional=True)],
    outputs=[Tensor("output", ("total_tokens", "num_v_heads", "head_dim_v"), dtype_from="q"),
             Tensor("final_state", ("num_seqs", "num_v_heads", "head_dim_v", "head_dim_k"), dtype="float32")],
    reference=_gdn_prefill_reference, init=_gdn_prefill_init, tags=("gdn", "prefill"), constraints=("len_cu == num_seqs + 1",),
    description="Gated delta rule over packed sequences (token-sequential or chunk-parallel WY execution, same result)", tolerance="bf16",
    test_sizes=_SIZES)
