"""Sampling templates (reference flashinfer/trace/templates/sampling.py).

Deterministic ops (softmax, renormalisation, masking) carry value references.  Samplers draw from a private Philox stream,
so their references return the *support*: a ``[batch, vocab]`` boolean mask of the tokens the filtered distribution can
emit (tolerance class ``support``: every sampled id must lie inside it)."""
import torch

from ..template import Const, Scalar, Tensor, TraceTemplate, Var

_AXES = [Var("batch_size"), Const("vocab_size", abbrev="v")]
_P = ("batch_size", "vocab_size")
_SIZES = {"vocab_size": 257}


def _probs(batch_size, vocab_size, device, seed, peaked=True):
    g = torch.Generator(device="cpu").manual_seed(seed)
    logits = torch.randn(batch_size, vocab_size, generator=g) * (3.0 if peaked else 1.0)
    return g, torch.softmax(logits, -1).to(device), logits.to(device)


# ---- deterministic
def _softmax_reference(logits, temperature=None):
    t = 1.0 if temperature is None else temperature
    t = t[:, None].to(torch.float32) if isinstance(t, torch.Tensor) else float(t)
    return torch.softmax(logits.to(torch.float32) / t, dim=-1)


def _softmax_init(*, batch_size=8, vocab_size=128256, device="cuda", seed=0):
    g, _, logits = _probs(batch_size, vocab_size, device, seed)
    return {"logits": logits, "temperature": (torch.rand(batch_size, generator=g) + 0.5).to(device)}


softmax_trace = TraceTemplate(
    op_type="sampling", name_fmt="softmax_v{vocab_size}", axes=_AXES,
    inputs=[Tensor("logits", _P), Tensor("temperature", ("batch_size",), optional=True, description="scalar or per row")],
    outputs=[Tensor("probs", _P, dtype="float32")], reference=_softmax_reference, init=_softmax_init, tags=("sampling",),
    description="Temperature-scaled safe softmax over the vocabulary", tolerance="fp32", test_sizes=_SIZES)


def _top_k_renorm_probs_reference(probs, top_k):
    p = probs.to(torch.float32)
    k = top_k if isinstance(top_k, torch.Tensor) else torch.full((p.shape[0],), int(top_k), device=p.device)
    kth = p.sort(-1, descending=True).values.gather(-1, (k.long().clamp(1, p.shape[-1]) - 1)[:, None])
    kept = torch.where(p >= kth, p, torch.zeros_like(p))        # ties with the k-th value are kept
    return kept / kept.sum(-1, keepdim=True)


def _top_k_init(*, batch_size=8, vocab_size=128256, device="cuda", seed=0):
    g, p, _ = _probs(batch_size, vocab_size, device, seed)
    return {"probs": p, "top_k": torch.randint(1, 50, (batch_size,), generator=g, dtype=torch.int32).to(device)}


top_k_renorm_probs_trace = TraceTemplate(
    op_type="sampling", name_fmt="top_k_renorm_probs_v{vocab_size}", axes=_AXES,
    inputs=[Tensor("probs", _P), Tensor("top_k", ("batch_size",), "int32", description="scalar or per row")],
    outputs=[Tensor("renorm_probs", _P, dtype="float32")], reference=_top_k_renorm_probs_reference, init=_top_k_init,
    tags=("sampling", "renorm"), description="Zero everything below the k-th largest probability and renormalise",
    tolerance="fp32", test_sizes=_SIZES)


def _top_p_renorm_probs_reference(probs, top_p):
    p = probs.to(torch.float32)
    sp, _ = p.sort(-1, descending=True)
    tp = top_p[:, None].to(torch.float32) if isinstance(top_p, torch.Tensor) else float(top_p)
    in_nucleus = (sp.cumsum(-1) - sp) < tp                       # a value is kept while the mass *before* it is < top_p
    thr = torch.where(in_nucleus, sp, torch.full_like(sp, float("inf"))).min(-1, keepdim=True).values
    kept = torch.where(p >= thr, p, torch.zeros_like(p))
    return kept / kept.sum(-1, keepdim=True)


def _top_p_init(*, batch_size=8, vocab_size=128256, device="cuda", seed=0):
    g, p, _ = _probs(batch_size, vocab_size, device, seed)
    return {"probs": p, "top_p": (torch.rand(batch_size, generator=g) * 0.8 + 0.1).to(device)}


top_p_renorm_probs_trace = TraceTemplate(
    op_type="sampling", name_fmt="top_p_renorm_probs_v{vocab_size}", axes=_AXES,
    inputs=[Tensor("probs", _P), Tensor("top_p", ("batch_size",), "float32", description="scalar or per row")],
    outputs=[Tensor("renorm_probs", _P, dtype="float32")], reference=_top_p_renorm_probs_reference, init=_top_p_init,
    tags=("sampling", "renorm"), description="Keep the smallest prefix of the sorted distribution reaching mass top_p",
    tolerance="fp32", test_sizes=_SIZES)


def _top_k_mask_logits_reference(logits, top_k):
    x = logits.to(torch.float32)
    k = top_k if isinstance(top_k, torch.Tensor) else torch.full((x.shape[0],), int(top_k), device=x.device)
    kth = x.sort(-1, descending=True).values.gather(-1, (k.long().clamp(1, x.shape[-1]) - 1)[:, None])
    return torch.where(x >= kth, x, torch.full_like(x, float("-inf")))


def _top_k_logits_init(*, batch_size=8, vocab_size=128256, device="cuda", seed=0):
    g, _, logits = _probs(batch_size, vocab_size, device, seed)
    return {"logits": logits, "top_k": torch.randint(1, 50, (batch_size,), generator=g, dtype=torch.int32).to(device)}


top_k_mask_logits_trace = TraceTemplate(
    op_type="sampling", name_fmt="top_k_mask_logits_v{vocab_size}", axes=_AXES,
    inputs=[Tensor("logits", _P), Tensor("top_k", ("batch_size",), "int32")],
    outputs=[Tensor("masked_logits", _P, dtype="float32")], reference=_top_k_mask_logits_reference, init=_top_k_logits_init,
    tags=("sampling", "mask"), description="-inf everywhere below the k-th largest logit", tolerance="fp32", test_sizes=_SIZES)


# ---- samplers: support masks
def _support_any(probs):
    return probs > 0


def _support_logits(logits):
    return torch.isfinite(logits)


def _support_top_k(probs, top_k):
    p = probs.to(torch.float32)
    k = top_k if isinstance(top_k, torch.Tensor) else torch.full((p.shape[0],), int(top_k), device=p.device)
    kth = p.sort(-1, descending=True).values.gather(-1, (k.long().clamp(1, p.shape[-1]) - 1)[:, None])
    return (p >= kth) & (p > 0)


def _support_top_p(probs, top_p):
    p = probs.to(torch.float32)
    sp, _ = p.sort(-1, descending=True)
    tp = top_p[:, None].to(torch.float32) if isinstance(top_p, torch.Tensor) else float(top_p)
    thr = torch.where((sp.cumsum(-1) - sp) < tp, sp, torch.full_like(sp, float("inf"))).min(-1, keepdim=True).values
    return (p >= thr) & (p > 0)


def _support_min_p(probs, min_p):
    p = probs.to(torch.float32)
    m = min_p[:, None].to(torch.float32) if isinstance(min_p, torch.Tensor) else float(min_p)
    return (p >= p.max(-1, keepdim=True).values * m) & (p > 0)


def _support_top_k_top_p(probs, top_k, top_p):
    """top_k_first: top-p is evaluated on the top-k-renormalised distribution."""
    p = probs.to(torch.float32)
    k = top_k if isinstance(top_k, torch.Tensor) else torch.full((p.shape[0],), int(top_k), device=p.device)
    kth = p.sort(-1, descending=True).values.gather(-1, (k.long().clamp(1, p.shape[-1]) - 1)[:, None])
    p = torch.where(p >= kth, p, torch.zeros_like(p))
    p = p / p.sum(-1, keepdim=True)
    sp, _ = p.sort(-1, descending=True)
    tp = top_p[:, None].to(torch.float32) if isinstance(top_p, torch.Tensor) else float(top_p)
    thr = torch.where((sp.cumsum(-1) - sp) < tp, sp, torch.full_like(sp, float("inf"))).min(-1, keepdim=True).values
    return (p >= thr) & (p > 0)


def _support_top_k_top_p_logits(logits, top_k, top_p):
    p = torch.softmax(logits.to(torch.float32), -1)
    k = top_k if isinstance(top_k, torch.Tensor) else torch.full((p.shape[0],), int(top_k), device=p.device)
    kth = p.sort(-1, descending=True).values.gather(-1, (k.long().clamp(1, p.shape[-1]) - 1)[:, None])
    p = torch.where(p >= kth, p, torch.zeros_like(p))
    p = p / p.sum(-1, keepdim=True)
    sp, _ = p.sort(-1, descending=True)
    tp = top_p[:, None].to(torch.float32) if isinstance(top_p, torch.Tensor) else float(top_p)
    thr = torch.where((sp.cumsum(-1) - sp) < tp, sp, torch.full_like(sp, float("inf"))).min(-1, keepdim=True).values
    return (p >= thr) & (p > 0)


def _probs_init(*, batch_size=8, vocab_size=128256, device="cuda", seed=0):
    return {"probs": _probs(batch_size, vocab_size, device, seed)[1]}


def _logits_init(*, batch_size=8, vocab_size=128256, device="cuda", seed=0):
    return {"logits": _probs(batch_size, vocab_size, device, seed)[2]}


def _min_p_init(*, batch_size=8, vocab_size=128256, device="cuda", seed=0):
    g, p, _ = _probs(batch_size, vocab_size, device, seed)
    return {"probs": p, "min_p": (torch.rand(batch_size, generator=g) * 0.3 + 0.05).to(device)}


def _top_k_top_p_init(*, batch_size=8, vocab_size=128256, device="cuda", seed=0):
    kw = _top_k_init(batch_size=batch_size, vocab_size=vocab_size, device=device, seed=seed)
    kw["top_p"] = _top_p_init(batch_size=batch_size, vocab_size=vocab_size, device=device, seed=seed + 1)["top_p"]
    return kw


def _top_k_top_p_logits_init(*, batch_size=8, vocab_size=128256, device="cuda", seed=0):
    kw = _top_k_top_p_init(batch_size=batch_size, vocab_size=vocab_size, device=device, seed=seed)
    kw["logits"] = _probs(batch_size, vocab_size, device, seed)[2]
    del kw["probs"]
    return kw


def _sampler(name, src, extra, ref, init, desc):
    return TraceTemplate(
        op_type="sampling", name_fmt=name + "_v{vocab_size}", axes=_AXES, inputs=[Tensor(src, _P)] + extra,
        outputs=[Tensor("samples", ("batch_size",), dtype="int32")], reference=ref, init=init, tags=("sampling", "stochastic"),
        description=desc + " (reference = support mask of the filtered distribution)", tolerance="support", test_sizes=_SIZES)


sampling_from_probs_trace = _sampler("sampling_from_probs", "probs", [], _support_any, _probs_init, "Inverse-CDF categorical sampling")
sampling_from_logits_trace = _sampler("sampling_from_logits", "logits", [], _support_logits, _logits_init,
                                      "Categorical sampling from logits")
top_k_sampling_from_probs_trace = _sampler("top_k_sampling_from_probs", "probs", [Tensor("top_k", ("batch_size",), "int32")],
                                           _support_top_k, _top_k_init, "Sorting-free rejection sampling restricted to the top-k")
top_p_sampling_from_probs_trace = _sampler("top_p_sampling_from_probs", "probs", [Tensor("top_p", ("batch_size",), "float32")],
                                           _support_top_p, _top_p_init, "Sorting-free nucleus sampling")
min_p_sampling_from_probs_trace = _sampler("min_p_sampling_from_probs", "probs", [Tensor("min_p", ("batch_size",), "float32")],
                                           _support_min_p, _min_p_init, "min-p sampling: keep p >= min_p * max(p)")
top_k_top_p_sampling_from_probs_trace = _sampler(
    "top_k_top_p_sampling_from_probs", "probs", [Tensor("top_k", ("batch_size",), "int32"), Tensor("top_p", ("batch_size",), "float32")],
    _support_top_k_top_p, _top_k_top_p_init, "top-k then top-p sampling")
top_k_top_p_sampling_from_logits_trace = _sampler(
    "top_k_top_p_sampling_from_logits", "logits", [Tensor("top_k", ("batch_size",), "int32"), Tensor("top_p", ("batch_size",), "float32")],
    _support_top_k_top_p_logits, _top_k_top_p_logits_init, "top-k mask, softmax, then top-p sampling")


# ---- speculative decoding
def _chain_speculative_sampling_reference(draft_probs, draft_token_ids, target_probs):
    """Accept draft token i with probability min(1, q_i / p_i); at the first rejection emit one token from
    normalise(relu(q - p)) and stop; if all are accepted emit a bonus token from the last target row.  Unused slots are -1.
    (Draws come from torch's global generator here: only distributions with a forced outcome compare exactly.)"""
    b, n, _ = draft_probs.shape
    out = torch.full((b, n + 1), -1, dtype=torch.int32, device=draft_probs.device)
    for i in range(b):
        pos = n
        for j in range(n):
            tok = int(draft_token_ids[i, j])
            q, p = float(target_probs[i, j, tok]), float(draft_probs[i, j, tok])
            if float(torch.rand(())) * p < q:
                out[i, j] = tok
            else:
                pos = j
                break
        dist = target_probs[i, pos].to(torch.float32)
        if pos < n:
            dist = (dist - draft_probs[i, pos].to(torch.float32)).clamp(min=0)
        out[i, pos] = int(torch.multinomial(dist / dist.sum(), 1))
    return out


def _chain_init(*, batch_size=4, num_speculate_tokens=3, vocab_size=32000, device="cuda", seed=0):
    """One-hot draft and target rows: rows 0, 2, ... agree everywhere (all accepted + bonus token), odd rows disagree at
    position 1 (rejected there, the residual distribution is one-hot as well) - the outcome does not depend on the draws."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    n = num_speculate_tokens
    draft_ids = torch.randint(0, vocab_size, (batch_size, n), generator=g)
    target_ids = torch.cat([draft_ids, torch.randint(0, vocab_size, (batch_size, 1), generator=g)], dim=1)
    for i in range(1, batch_size, 2):
        target_ids[i, min(1, n - 1)] = (draft_ids[i, min(1, n - 1)] + 1) % vocab_size
    draft = torch.zeros(batch_size, n, vocab_size).scatter_(2, draft_ids[..., None], 1.0)
    target = torch.zeros(batch_size, n + 1, vocab_size).scatter_(2, target_ids[..., None], 1.0)
    return {"draft_probs": draft.to(device), "draft_token_ids": draft_ids.int().to(device), "target_probs": target.to(device)}


chain_speculative_sampling_trace = TraceTemplate(
    op_type="sampling", name_fmt="chain_speculative_sampling_n{num_speculate_tokens}_v{vocab_size}",
    axes=[Var("batch_size"), Const("num_speculate_tokens", abbrev="n"), Const("vocab_size", abbrev="v")],
    inputs=[Tensor("draft_probs", ("batch_size", "num_speculate_tokens", "vocab_size")),
            Tensor("draft_token_ids", ("batch_size", "num_speculate_tokens"), "int32"),
            Tensor("target_probs", ("batch_size", "num_speculate_plus_one", "vocab_size"))],
    outputs=[Tensor("output_token_ids", ("batch_size", "num_speculate_plus_one"), dtype="int32")],
    reference=_chain_speculative_sampling_reference, init=_chain_init, tags=("sampling", "speculative"),
    constraints=("num_speculate_plus_one == num_speculate_tokens + 1",),
    description="Chain speculative-decoding verification (accept / reject / resample)", tolerance="exact",
    test_sizes={"num_speculate_tokens": 3, "vocab_size": 97})
