"""Templates for ops outside the big categories: radix top-k, MLA key assembly, Mamba decode step
(reference flashinfer/trace/templates/{sampling,attention,mamba}.py hold their counterparts)."""
import torch

from ..template import Const, Scalar, Tensor, TraceTemplate, Var


def _top_k_reference(input, k):
    v, i = torch.topk(input.to(torch.float32), k, dim=-1, sorted=True)
    return v.to(input.dtype), i


def _top_k_init(*, batch_size=8, vocab_size=128256, k=64, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return {"input": torch.randn(batch_size, vocab_size, generator=g).to(device), "k": k, "sorted": True}


def _top_k_compare(got, expected, kwargs):
    (v, i), (v_ref, _) = got, expected
    assert torch.equal(v.float(), v_ref.float()), "top-k values differ"
    assert torch.equal(kwargs["input"].gather(-1, i.long()).float(), v.float()), "indices do not address the returned values"


top_k_trace = TraceTemplate(
    op_type="topk", name_fmt="top_k_v{vocab_size}_k{k}", axes=[Var("batch_size"), Const("vocab_size", abbrev="v"), Const("k")],
    inputs=[Tensor("input", ("batch_size", "vocab_size")), Scalar("k", "int32")],
    outputs=[Tensor("values", ("batch_size", "k"), dtype_from="input"), Tensor("indices", ("batch_size", "k"), dtype="int64")],
    reference=_top_k_reference, init=_top_k_init, compare=_top_k_compare, tags=("topk",),
    description="Row-wise top-k (radix select); values sorted descending when sorted=True", test_sizes={"vocab_size": 513, "k": 7})


def _concat_mla_k_reference(k_nope, k_rope):
    """k[t, h] = [k_nope[t, h] | k_rope[t, 0]]: the single rope head is broadcast to every head."""
    return torch.cat([k_nope, k_rope.expand(-1, k_nope.shape[1], -1)], dim=-1)


def _concat_mla_k_init(*, num_tokens=64, num_heads=128, nope_dim=128, rope_dim=64, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    mk = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16).to(device)  # noqa: E731
    return {"k": torch.empty(num_tokens, num_heads, nope_dim + rope_dim, dtype=torch.bfloat16, device=device),
            "k_nope": mk(num_tokens, num_heads, nope_dim), "k_rope": mk(num_tokens, 1, rope_dim)}


concat_mla_k_trace = TraceTemplate(
    op_type="concat", name_fmt="concat_mla_k_h{num_heads}_n{nope_dim}_r{rope_dim}",
    axes=[Var("num_tokens"), Const("num_heads", abbrev="h"), Const("nope_dim", abbrev="n"), Const("rope_dim", abbrev="r")],
    inputs=[Tensor("k_nope", ("num_tokens", "num_heads", "nope_dim")), Tensor("k_rope", ("num_tokens", "one", "rope_dim"))],
    outputs=[Tensor("k", ("num_tokens", "num_heads", "qk_dim"), dtype_from="k_nope", param="k")], reference=_concat_mla_k_reference,
    init=_concat_mla_k_init, tags=("mla", "memory"), constraints=("one == 1", "qk_dim == nope_dim + rope_dim"),
    description="Assemble MLA prefill keys from per-head no-rope parts and the shared rope part", tolerance="exact",
    test_sizes={"num_heads": 4, "nope_dim": 32, "rope_dim": 16})


def _selective_state_update_reference(state, x, dt, A, B, C, D=None, z=None, dt_bias=None, dt_softplus=False):
    """One Mamba-2 decode step per (batch, head): state [b, h, dim, n]; x, dt, z [b, h, dim]; A [h, dim, n]; B, C [b, g, n]; D [h, dim].
    state' = state * exp(dt A) + dt x B;  y = state' . C + D x;  y *= silu(z).  Returns (y, state')."""
    b, h, dim, n = state.shape
    dtf = dt.to(torch.float32) + (dt_bias.to(torch.float32) if dt_bias is not None else 0.0)
    if dt_softplus:
        dtf = torch.nn.functional.softplus(dtf)
    rep = h // B.shape[1]
    Bf = B.to(torch.float32).repeat_interleave(rep, dim=1)
    Cf = C.to(torch.float32).repeat_interleave(rep, dim=1)
    dA = torch.exp(dtf[..., None] * A.to(torch.float32))
    new = state.to(torch.float32) * dA + (dtf * x.to(torch.float32))[..., None] * Bf[:, :, None, :]
    y = (new * Cf[:, :, None, :]).sum(-1)
    if D is not None:
        y = y + D.to(torch.float32) * x.to(torch.float32)
    if z is not None:
        y = y * torch.nn.functional.silu(z.to(torch.float32))
    return y.to(x.dtype), new.to(state.dtype)


def _ssu_init(*, batch_size=8, nheads=64, dim=64, dstate=128, ngroups=8, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    return {"state": r(batch_size, nheads, dim, dstate).to(device), "x": r(batch_size, nheads, dim).to(torch.bfloat16).to(device),
            "dt": (r(batch_size, nheads, dim) * 0.5).to(torch.bfloat16).to(device),
            "A": (-torch.rand(nheads, dim, dstate, generator=g) - 0.1).to(device), "B": r(batch_size, ngroups, dstate).to(torch.bfloat16).to(device),
            "C": r(batch_size, ngroups, dstate).to(torch.bfloat16).to(device), "D": r(nheads, dim).to(device),
            "z": r(batch_size, nheads, dim).to(torch.bfloat16).to(device), "dt_bias": (r(nheads, dim) * 0.1).to(device), "dt_softplus": True}


selective_state_update_trace = TraceTemplate(
    op_type="mamba", name_fmt="selective_state_update_h{nheads}_d{dim}_n{dstate}_g{ngroups}",
    axes=[Var("batch_size"), Const("nheads", abbrev="h"), Const("dim", abbrev="d"), Const("dstate", abbrev="n"), Const("ngroups", abbrev="g")],
    inputs=[Tensor("state", ("batch_size", "nheads", "dim", "dstate")), Tensor("x", ("batch_size", "nheads", "dim")),
            Tensor("dt", ("batch_size", "nheads", "dim")), Tensor("A", ("nheads", "dim", "dstate")), Tensor("B", ("batch_size", "ngroups", "dstate")),
            Tensor("C", ("batch_size", "ngroups", "dstate")), Tensor("D", ("nheads", "dim"), optional=True),
            Tensor("z", ("batch_size", "nheads", "dim"), optional=True), Tensor("dt_bias", ("nheads", "dim"), optional=True),
            Scalar("dt_softplus", "bool", optional=True)],
    outputs=[Tensor("y", ("batch_size", "nheads", "dim"), dtype_from="x"),
             Tensor("state_out", ("batch_size", "nheads", "dim", "dstate"), dtype_from="state", param="state")],
    reference=_selective_state_update_reference, init=_ssu_init, tags=("mamba", "decode", "inplace"),
    description="Mamba-2 selective state-space decode step (state updated in place)", tolerance="bf16",
    test_sizes={"nheads": 4, "dim": 8, "dstate": 16, "ngroups": 2})
