"""Attention-state merge templates (reference flashinfer/trace/templates/cascade.py).  A state is (v, s): the normalised
partial output and the log-sum-exp of its scaled logits **in base 2** (s = log2 sum_j 2^(logit_j * log2 e)), the convention
of every ``return_lse`` in this API.  merge(a, b) weights the outputs by 2^s_a, 2^s_b; s = log2(2^s_a + 2^s_b)."""
import torch

from ..template import Const, Tensor, TraceTemplate, Var

_AXES = [Var("seq_len"), Const("num_heads", abbrev="h"), Const("head_dim", abbrev="d")]
_V, _S = ("seq_len", "num_heads", "head_dim"), ("seq_len", "num_heads")
_SIZES = {"num_heads": 4, "head_dim": 64}


def _merge_state_reference(v_a, s_a, v_b, s_b):
    sa, sb = s_a.to(torch.float32), s_b.to(torch.float32)
    m = torch.maximum(sa, sb)
    wa, wb = torch.exp2(sa - m), torch.exp2(sb - m)
    v = (v_a.to(torch.float32) * wa[..., None] + v_b.to(torch.float32) * wb[..., None]) / (wa + wb)[..., None]
    return v.to(v_a.dtype), m + torch.log2(wa + wb)


def _merge_state_init(*, seq_len=16, num_heads=32, head_dim=128, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    mk = lambda: torch.randn(seq_len, num_heads, head_dim, generator=g).to(torch.bfloat16).to(device)  # noqa: E731
    ms = lambda: (torch.randn(seq_len, num_heads, generator=g) * 3).to(device)  # noqa: E731
    return {"v_a": mk(), "s_a": ms(), "v_b": mk(), "s_b": ms()}


merge_state_trace = TraceTemplate(
    op_type="cascade", name_fmt="merge_state_h{num_heads}_d{head_dim}", axes=_AXES,
    inputs=[Tensor("v_a", _V), Tensor("s_a", _S, "float32"), Tensor("v_b", _V), Tensor("s_b", _S, "float32")],
    outputs=[Tensor("v", _V, dtype_from="v_a"), Tensor("s", _S, dtype="float32")], reference=_merge_state_reference,
    init=_merge_state_init, tags=("cascade",), description="Merge two partial attention states (LSE in base 2)",
    tolerance="bf16", test_sizes=_SIZES)


def _merge_state_in_place_reference(v, s, v_other, s_other, mask=None):
    sa, sb = s.to(torch.float32), s_other.to(torch.float32)
    m = torch.maximum(sa, sb)
    wa, wb = torch.exp2(sa - m), torch.exp2(sb - m)
    vm = ((v.to(torch.float32) * wa[..., None] + v_other.to(torch.float32) * wb[..., None]) / (wa + wb)[..., None]).to(v.dtype)
    sm = m + torch.log2(wa + wb)
    if mask is not None:
        keep = ~mask.bool()
        vm = torch.where(keep[:, None, None], v, vm)
        sm = torch.where(keep[:, None], sa, sm)
    return vm, sm


def _merge_state_in_place_init(*, seq_len=16, num_heads=32, head_dim=128, device="cuda", seed=0):
    kw = _merge_state_init(seq_len=seq_len, num_heads=num_heads, head_dim=head_dim, device=device, seed=seed)
    g = torch.Generator(device="cpu").manual_seed(seed + 7)
    return {"v": kw["v_a"], "s": kw["s_a"], "v_other": kw["v_b"], "s_other": kw["s_b"],
            "mask": (torch.rand(seq_len, generator=g) > 0.3).to(device)}


merge_state_in_place_trace = TraceTemplate(
    op_type="cascade", name_fmt="merge_state_in_place_h{num_heads}_d{head_dim}", axes=_AXES,
    inputs=[Tensor("v", _V), Tensor("s", _S, "float32"), Tensor("v_other", _V), Tensor("s_other", _S, "float32"),
            Tensor("mask", ("seq_len",), "bool", optional=True, description="rows with False keep (v, s) untouched")],
    outputs=[Tensor("v_out", _V, dtype_from="v", param="v"), Tensor("s_out", _S, dtype="float32", param="s")],
    reference=_merge_state_in_place_reference, init=_merge_state_in_place_init, tags=("cascade", "inplace"),
    description="Merge (v_other, s_other) into (v, s) in place, optionally only on masked rows", tolerance="bf16", test_sizes=_SIZES)


def _merge_states_reference(v, s):
    sf = s.to(torch.float32)
    m = sf.max(dim=1, keepdim=True).values
    w = torch.exp2(sf - m)
    out = (v.to(torch.float32) * w[..., None]).sum(1) / w.sum(1)[..., None]
    return out.to(v.dtype), (m + torch.log2(w.sum(1, keepdim=True))).squeeze(1)


def _merge_states_init(*, seq_len=16, num_states=4, num_heads=32, head_dim=128, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return {"v": torch.randn(seq_len, num_states, num_heads, head_dim, generator=g).to(torch.bfloat16).to(device),
            "s": (torch.randn(seq_len, num_states, num_heads, generator=g) * 3).to(device)}


merge_states_trace = TraceTemplate(
    op_type="cascade", name_fmt="merge_states_h{num_heads}_d{head_dim}",
    axes=[Var("seq_len"), Var("num_states"), Const("num_heads", abbrev="h"), Const("head_dim", abbrev="d")],
    inputs=[Tensor("v", ("seq_len", "num_states", "num_heads", "head_dim")), Tensor("s", ("seq_len", "num_states", "num_heads"), "float32")],
    outputs=[Tensor("v_merged", _V, dtype_from="v"), Tensor("s_merged", _S, dtype="float32")], reference=_merge_states_reference,
    init=_merge_states_init, tags=("cascade",), description="Merge num_states partial attention states per row", tolerance="bf16",
    test_sizes=_SIZES)
