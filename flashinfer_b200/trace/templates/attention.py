"""Attention templates (reference flashinfer/trace/templates/attention.py).  GQA: query head h reads KV head
h // (num_qo_heads / num_kv_heads).  LSE is base 2 (see cascade.py).  Wrapper ``run`` methods read the page table the
wrapper captured in ``plan()`` (``param="self.<attr>"``)."""
import math

import torch

from ..template import Const, Scalar, Tensor, TraceTemplate, Var

_H = [Const("num_qo_heads", abbrev="h"), Const("num_kv_heads", abbrev="kv"), Const("head_dim", abbrev="d")]
_SIZES = {"num_qo_heads": 4, "num_kv_heads": 2, "head_dim": 64, "page_size": 4}


# ------------------------------------------------------------------ single request
def _single_decode_reference(q, k, v, sm_scale=None):
    """q [H, D]; k, v [L, Hkv, D] (NHD)."""
    h, d = q.shape
    group = h // k.shape[1]
    scale = sm_scale if sm_scale is not None else 1.0 / (d ** 0.5)
    kf = k.to(torch.float32).repeat_interleave(group, dim=1)
    vf = v.to(torch.float32).repeat_interleave(group, dim=1)
    logits = torch.einsum("hd,lhd->hl", q.to(torch.float32), kf) * scale
    return torch.einsum("hl,lhd->hd", torch.softmax(logits, -1), vf).to(q.dtype)


def _single_decode_init(*, kv_len=1024, num_qo_heads=32, num_kv_heads=8, head_dim=128, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    mk = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16).to(device)  # noqa: E731
    return {"q": mk(num_qo_heads, head_dim), "k": mk(kv_len, num_kv_heads, head_dim), "v": mk(kv_len, num_kv_heads, head_dim)}


single_decode_with_kv_cache_trace = TraceTemplate(
    op_type="gqa_single", name_fmt="single_decode_h{num_qo_heads}_kv{num_kv_heads}_d{head_dim}", axes=[Var("kv_len")] + _H,
    inputs=[Tensor("q", ("num_qo_heads", "head_dim")), Tensor("k", ("kv_len", "num_kv_heads", "head_dim")),
            Tensor("v", ("kv_len", "num_kv_heads", "head_dim")), Scalar("sm_scale", optional=True)],
    outputs=[Tensor("output", ("num_qo_heads", "head_dim"), dtype_from="q")], reference=_single_decode_reference,
    init=_single_decode_init, tags=("attention", "decode"), description="One query token against a contiguous NHD KV cache",
    tolerance="bf16", test_sizes=_SIZES)


def _single_prefill_reference(q, k, v, causal=False, sm_scale=None):
    """q [Lq, H, D]; k, v [Lk, Hkv, D]; causal aligns the last query with the last key."""
    lq, h, d = q.shape
    lk = k.shape[0]
    group = h // k.shape[1]
    scale = sm_scale if sm_scale is not None else 1.0 / (d ** 0.5)
    kf = k.to(torch.float32).repeat_interleave(group, dim=1)
    vf = v.to(torch.float32).repeat_interleave(group, dim=1)
    logits = torch.einsum("qhd,khd->hqk", q.to(torch.float32), kf) * scale
    if causal:
        qi = torch.arange(lq, device=q.device)[:, None] + (lk - lq)
        logits = logits.masked_fill(torch.arange(lk, device=q.device)[None, :] > qi, float("-inf"))
    return torch.einsum("hqk,khd->qhd", torch.softmax(logits, -1), vf).to(q.dtype)


def _single_prefill_init(*, qo_len=128, kv_len=256, num_qo_heads=32, num_kv_heads=8, head_dim=128, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    mk = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16).to(device)  # noqa: E731
    return {"q": mk(qo_len, num_qo_heads, head_dim), "k": mk(kv_len, num_kv_heads, head_dim), "v": mk(kv_len, num_kv_heads, head_dim),
            "causal": True}


single_prefill_with_kv_cache_trace = TraceTemplate(
    op_type="gqa_single", name_fmt="single_prefill_h{num_qo_heads}_kv{num_kv_heads}_d{head_dim}", axes=[Var("qo_len"), Var("kv_len")] + _H,
    inputs=[Tensor("q", ("qo_len", "num_qo_heads", "head_dim")), Tensor("k", ("kv_len", "num_kv_heads", "head_dim")),
            Tensor("v", ("kv_len", "num_kv_heads", "head_dim")), Scalar("causal", "bool", optional=True), Scalar("sm_scale", optional=True)],
    outputs=[Tensor("output", ("qo_len", "num_qo_heads", "head_dim"), dtype_from="q")], reference=_single_prefill_reference,
    init=_single_prefill_init, tags=("attention", "prefill"), description="One request, contiguous NHD q / k / v, optional causal mask",
    tolerance="bf16", test_sizes=dict(_SIZES, qo_len=7, kv_len=19))


# ------------------------------------------------------------------ batched, paged
def _paged_tables(lens, page_size, device, g):
    pages_per = [(n + page_size - 1) // page_size for n in lens]
    total = sum(pages_per)
    indices = torch.randperm(total + 2, generator=g)[:total].int()
    indptr = torch.tensor([0] + list(torch.tensor(pages_per).cumsum(0)), dtype=torch.int32)
    last = torch.tensor([(n - 1) % page_size + 1 for n in lens], dtype=torch.int32)
    return indices.to(device), indptr.to(device), last.to(device), total + 2


def _gqa_paged_decode_reference(q, k_cache, v_cache, kv_indptr, kv_indices, kv_last_page_len, sm_scale=None):
    """k_cache / v_cache [pages, page_size, Hkv, D] (NHD).  Returns (output, lse base 2)."""
    b, h, d = q.shape
    page_size = k_cache.shape[1]
    group = h // k_cache.shape[2]
    scale = sm_scale if sm_scale is not None else 1.0 / (d ** 0.5)
    out = torch.zeros(b, h, d, dtype=torch.float32, device=q.device)
    lse = torch.zeros(b, h, dtype=torch.float32, device=q.device)
    for i in range(b):
        pages = kv_indices[int(kv_indptr[i]): int(kv_indptr[i + 1])].long()
        n = (pages.numel() - 1) * page_size + int(kv_last_page_len[i])
        k = k_cache[pages].reshape(-1, *k_cache.shape[2:])[:n].to(torch.float32).repeat_interleave(group, dim=1)
        v = v_cache[pages].reshape(-1, *v_cache.shape[2:])[:n].to(torch.float32).repeat_interleave(group, dim=1)
        logits = torch.einsum("hd,lhd->hl", q[i].to(torch.float32), k) * scale
        out[i] = torch.einsum("hl,lhd->hd", torch.softmax(logits, -1), v)
        lse[i] = torch.logsumexp(logits, -1) * 1.4426950408889634
    return out.to(q.dtype), lse


def _gqa_paged_decode_init(*, batch_size=8, num_qo_heads=32, num_kv_heads=8, head_dim=128, page_size=16, kv_len=None, device="cuda", seed=0):
    """``kv_len=None``: ragged lengths in [1, 6 pages); an integer gives every request that many cached tokens."""
    import flashinfer_b200 as fi

    g = torch.Generator(device="cpu").manual_seed(seed)
    lens = [int(x) for x in torch.randint(1, 6 * page_size, (batch_size,), generator=g)] if kv_len is None else [int(kv_len)] * batch_size
    indices, indptr, last, num_pages = _paged_tables(lens, page_size, device, g)
    mk = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16).to(device)  # noqa: E731
    w = fi.BatchDecodeWithPagedKVCacheWrapper(torch.empty(32 << 20, dtype=torch.uint8, device=device), "NHD")
    w.plan(indptr, indices, last, num_qo_heads, num_kv_heads, head_dim, page_size, q_data_type=torch.bfloat16)
    return {"self": w, "q": mk(batch_size, num_qo_heads, head_dim),
            "paged_kv_cache": (mk(num_pages, page_size, num_kv_heads, head_dim), mk(num_pages, page_size, num_kv_heads, head_dim)),
            "return_lse": True}


_PAGED = ("num_pages", "page_size", "num_kv_heads", "head_dim")
gqa_paged_decode_trace = TraceTemplate(
    op_type="gqa_paged", name_fmt="gqa_paged_decode_h{num_qo_heads}_kv{num_kv_heads}_d{head_dim}_ps{page_size}",
    axes=[Var("batch_size"), Var("num_pages"), Var("len_indptr"), Var("num_kv_indices")] + _H + [Const("page_size", abbrev="ps")],
    inputs=[Tensor("q", ("batch_size", "num_qo_heads", "head_dim")),
            Tensor("k_cache", _PAGED, param="paged_kv_cache", tuple_idx=0), Tensor("v_cache", _PAGED, param="paged_kv_cache", tuple_idx=1),
            Tensor("kv_indptr", ("len_indptr",), "int32", param="self._kv_indptr_host"),
            Tensor("kv_indices", ("num_kv_indices",), "int32", param="self._kv_indices"),
            Tensor("kv_last_page_len", ("batch_size",), "int32", param="self._kv_last_host"),
            Scalar("sm_scale", param="self._sm_scale")],
    outputs=[Tensor("output", ("batch_size", "num_qo_heads", "head_dim"), dtype_from="q"),
             Tensor("lse", ("batch_size", "num_qo_heads"), dtype="float32")],
    reference=_gqa_paged_decode_reference, init=_gqa_paged_decode_init, tags=("attention", "decode", "paged"),
    constraints=("len_indptr == batch_size + 1",), description="Batched GQA decode over a paged KV cache (NHD pages)",
    tolerance="bf16", test_sizes=_SIZES)


def _gqa_paged_prefill_reference(q, k_cache, v_cache, qo_indptr, kv_indptr, kv_indices, kv_last_page_len, causal=True, sm_scale=None):
    h, d = q.shape[1:]
    page_size = k_cache.shape[1]
    group = h // k_cache.shape[2]
    scale = sm_scale if sm_scale is not None else 1.0 / (d ** 0.5)
    out = torch.zeros(q.shape, dtype=torch.float32, device=q.device)
    lse = torch.zeros(q.shape[0], h, dtype=torch.float32, device=q.device)
    for i in range(qo_indptr.numel() - 1):
        qs, qe = int(qo_indptr[i]), int(qo_indptr[i + 1])
        pages = kv_indices[int(kv_indptr[i]): int(kv_indptr[i + 1])].long()
        n = (pages.numel() - 1) * page_size + int(kv_last_page_len[i])
        k = k_cache[pages].reshape(-1, *k_cache.shape[2:])[:n].to(torch.float32).repeat_interleave(group, dim=1)
        v = v_cache[pages].reshape(-1, *v_cache.shape[2:])[:n].to(torch.float32).repeat_interleave(group, dim=1)
        logits = torch.einsum("qhd,khd->hqk", q[qs:qe].to(torch.float32), k) * scale
        if causal:
            qi = torch.arange(qe - qs, device=q.device)[:, None] + (n - (qe - qs))
            logits = logits.masked_fill(torch.arange(n, device=q.device)[None, :] > qi, float("-inf"))
        out[qs:qe] = torch.einsum("hqk,khd->qhd", torch.softmax(logits, -1), v)
        lse[qs:qe] = (torch.logsumexp(logits, -1) * 1.4426950408889634).transpose(0, 1)
    return out.to(q.dtype), lse


def _gqa_paged_prefill_init(*, batch_size=4, num_qo_heads=32, num_kv_heads=8, head_dim=128, page_size=16, qo_len=None, kv_len=None,
                            device="cuda", seed=0):
    """``qo_len`` / ``kv_len`` None: ragged; integers give every request that many new / total tokens."""
    import flashinfer_b200 as fi

    g = torch.Generator(device="cpu").manual_seed(seed)
    q_lens = [int(x) for x in torch.randint(1, 3 * page_size, (batch_size,), generator=g)] if qo_len is None else [int(qo_len)] * batch_size
    kv_lens = [ql + int(x) for ql, x in zip(q_lens, torch.randint(0, 4 * page_size, (batch_size,), generator=g))] if kv_len is None \
        else [max(int(kv_len), ql) for ql in q_lens]
    indices, indptr, last, num_pages = _paged_tables(kv_lens, page_size, device, g)
    qo_indptr = torch.tensor([0] + list(torch.tensor(q_lens).cumsum(0)), dtype=torch.int32, device=device)
    mk = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16).to(device)  # noqa: E731
    w = fi.BatchPrefillWithPagedKVCacheWrapper(torch.empty(32 << 20, dtype=torch.uint8, device=device), "NHD")
    w.plan(qo_indptr, indptr, indices, last, num_qo_heads, num_kv_heads, head_dim, page_size, causal=True, q_data_type=torch.bfloat16)
    return {"self": w, "q": mk(sum(q_lens), num_qo_heads, head_dim),
            "paged_kv_cache": (mk(num_pages, page_size, num_kv_heads, head_dim), mk(num_pages, page_size, num_kv_heads, head_dim)),
            "return_lse": True}


gqa_paged_prefill_trace = TraceTemplate(
    op_type="gqa_paged", name_fmt="gqa_paged_prefill_causal_h{num_qo_heads}_kv{num_kv_heads}_d{head_dim}_ps{page_size}",
    axes=[Var("total_q"), Var("batch_size"), Var("num_pages"), Var("len_indptr"), Var("num_kv_indices")] + _H + [Const("page_size", abbrev="ps")],
    inputs=[Tensor("q", ("total_q", "num_qo_heads", "head_dim")),
            Tensor("k_cache", _PAGED, param="paged_kv_cache", tuple_idx=0), Tensor("v_cache", _PAGED, param="paged_kv_cache", tuple_idx=1),
            Tensor("qo_indptr", ("len_indptr",), "int32", param="self._qo_indptr_host"),
            Tensor("kv_indptr", ("len_indptr",), "int32", param="self._kv_indptr_host"),
            Tensor("kv_indices", ("num_kv_indices",), "int32", param="self._kv_indices"),
            Tensor("kv_last_page_len", ("batch_size",), "int32", param="self._kv_last_host"),
            Scalar("causal", "bool", param="self._causal"), Scalar("sm_scale", param="self._sm_scale")],
    outputs=[Tensor("output", ("total_q", "num_qo_heads", "head_dim"), dtype_from="q"), Tensor("lse", ("total_q", "num_qo_heads"), dtype="float32")],
    reference=_gqa_paged_prefill_reference, init=_gqa_paged_prefill_init, tags=("attention", "prefill", "paged"),
    constraints=("len_indptr == batch_size + 1",), description="Batched causal GQA prefill / append over a paged KV cache",
    tolerance="bf16", test_sizes=_SIZES)


def _gqa_ragged_prefill_reference(q, k, v, qo_indptr, kv_indptr, causal=True, sm_scale=None):
    h, d = q.shape[1:]
    group = h // k.shape[1]
    scale = sm_scale if sm_scale is not None else 1.0 / (d ** 0.5)
    out = torch.zeros(q.shape[0], h, v.shape[-1], dtype=torch.float32, device=q.device)
    lse = torch.zeros(q.shape[0], h, dtype=torch.float32, device=q.device)
    for i in range(qo_indptr.numel() - 1):
        qs, qe, ks, ke = int(qo_indptr[i]), int(qo_indptr[i + 1]), int(kv_indptr[i]), int(kv_indptr[i + 1])
        kf = k[ks:ke].to(torch.float32).repeat_interleave(group, dim=1)
        vf = v[ks:ke].to(torch.float32).repeat_interleave(group, dim=1)
        logits = torch.einsum("qhd,khd->hqk", q[qs:qe].to(torch.float32), kf) * scale
        if causal:
            qi = torch.arange(qe - qs, device=q.device)[:, None] + ((ke - ks) - (qe - qs))
            logits = logits.masked_fill(torch.arange(ke - ks, device=q.device)[None, :] > qi, float("-inf"))
        out[qs:qe] = torch.einsum("hqk,khd->qhd", torch.softmax(logits, -1), vf)
        lse[qs:qe] = (torch.logsumexp(logits, -1) * 1.4426950408889634).transpose(0, 1)
    return out.to(q.dtype), lse


def _gqa_ragged_prefill_init(*, batch_size=4, num_qo_heads=32, num_kv_heads=8, head_dim=128, seq_len=None, device="cuda", seed=0):
    import flashinfer_b200 as fi

    g = torch.Generator(device="cpu").manual_seed(seed)
    lens = [int(x) for x in torch.randint(1, 40, (batch_size,), generator=g)] if seq_len is None else [int(seq_len)] * batch_size
    indptr = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=device)
    mk = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16).to(device)  # noqa: E731
    w = fi.BatchPrefillWithRaggedKVCacheWrapper(torch.empty(32 << 20, dtype=torch.uint8, device=device), "NHD")
    w.plan(indptr, indptr, num_qo_heads, num_kv_heads, head_dim, causal=True, q_data_type=torch.bfloat16)
    n = sum(lens)
    return {"self": w, "q": mk(n, num_qo_heads, head_dim), "k": mk(n, num_kv_heads, head_dim), "v": mk(n, num_kv_heads, head_dim),
            "return_lse": True}


gqa_ragged_prefill_trace = TraceTemplate(
    op_type="gqa_ragged", name_fmt="gqa_ragged_prefill_causal_h{num_qo_heads}_kv{num_kv_heads}_d{head_dim}",
    axes=[Var("total_q"), Var("total_kv"), Var("len_indptr")] + _H,
    inputs=[Tensor("q", ("total_q", "num_qo_heads", "head_dim")), Tensor("k", ("total_kv", "num_kv_heads", "head_dim")),
            Tensor("v", ("total_kv", "num_kv_heads", "head_dim")), Tensor("qo_indptr", ("len_indptr",), "int32", param="self._qo_indptr_host"),
            Tensor("kv_indptr", ("len_indptr",), "int32", param="self._kv_indptr_ragged_host"),
            Scalar("causal", "bool", param="self._causal"), Scalar("sm_scale", param="self._sm_scale")],
    outputs=[Tensor("output", ("total_q", "num_qo_heads", "head_dim"), dtype_from="q"), Tensor("lse", ("total_q", "num_qo_heads"), dtype="float32")],
    reference=_gqa_ragged_prefill_reference, init=_gqa_ragged_prefill_init, tags=("attention", "prefill", "ragged"),
    description="Batched causal GQA prefill over ragged (packed) q / k / v", tolerance="bf16", test_sizes=_SIZES)


# ------------------------------------------------------------------ MLA
def _mla_paged_reference(q_nope, q_pe, ckv_cache, kpe_cache, qo_indptr, kv_indptr, kv_indices, kv_len_arr, causal=True, sm_scale=None):
    """Matrix-absorbed MLA: logits = q_nope . ckv + q_pe . kpe over the request's pages, output = softmax . ckv."""
    page_size = ckv_cache.shape[1]
    out = torch.zeros(q_nope.shape, dtype=torch.float32, device=q_nope.device)
    lse = torch.zeros(q_nope.shape[:2], dtype=torch.float32, device=q_nope.device)
    for i in range(qo_indptr.numel() - 1):
        qs, qe = int(qo_indptr[i]), int(qo_indptr[i + 1])
        n = int(kv_len_arr[i])
        pages = kv_indices[int(kv_indptr[i]): int(kv_indptr[i + 1])].long()
        ckv = ckv_cache[pages].reshape(-1, ckv_cache.shape[-1])[:n].to(torch.float32)
        kpe = kpe_cache[pages].reshape(-1, kpe_cache.shape[-1])[:n].to(torch.float32)
        logits = (torch.einsum("qhc,kc->hqk", q_nope[qs:qe].to(torch.float32), ckv)
                  + torch.einsum("qhr,kr->hqk", q_pe[qs:qe].to(torch.float32), kpe)) * sm_scale
        if causal:
            qi = torch.arange(qe - qs, device=logits.device)[:, None] + (n - (qe - qs))
            logits = logits.masked_fill(torch.arange(n, device=logits.device)[None, :] > qi, float("-inf"))
        out[qs:qe] = torch.einsum("hqk,kc->qhc", torch.softmax(logits, -1), ckv)
        lse[qs:qe] = (torch.logsumexp(logits, -1) * 1.4426950408889634).transpose(0, 1)
    return out.to(q_nope.dtype), lse


def _mla_paged_init(*, batch_size=4, num_heads=128, head_dim_ckv=512, head_dim_kpe=64, page_size=64, kv_len=None, device="cuda", seed=0):
    import flashinfer_b200 as fi

    g = torch.Generator(device="cpu").manual_seed(seed)
    kv_lens = [int(x) for x in torch.randint(1, 3 * page_size, (batch_size,), generator=g)] if kv_len is None else [int(kv_len)] * batch_size
    indices, indptr, _, num_pages = _paged_tables(kv_lens, page_size, device, g)
    qo_indptr = torch.arange(batch_size + 1, dtype=torch.int32, device=device)
    mk = lambda *s: (torch.randn(*s, generator=g) * 0.5).to(torch.bfloat16).to(device)  # noqa: E731
    w = fi.BatchMLAPagedAttentionWrapper(torch.empty(64 << 20, dtype=torch.uint8, device=device))
    scale = 1.0 / math.sqrt(128 + head_dim_kpe)
    w.plan(qo_indptr, indptr, indices, torch.tensor(kv_lens, dtype=torch.int32, device=device), num_heads, head_dim_ckv, head_dim_kpe,
           page_size, True, scale, torch.bfloat16, torch.bfloat16)
    return {"self": w, "q_nope": mk(batch_size, num_heads, head_dim_ckv), "q_pe": mk(batch_size, num_heads, head_dim_kpe),
            "ckv_cache": mk(num_pages, page_size, head_dim_ckv), "kpe_cache": mk(num_pages, page_size, head_dim_kpe), "return_lse": True}


mla_paged_trace = TraceTemplate(
    op_type="mla_paged", name_fmt="mla_paged_h{num_heads}_ckv{head_dim_ckv}_kpe{head_dim_kpe}_ps{page_size}",
    axes=[Var("total_q"), Var("batch_size"), Var("num_pages"), Var("len_indptr"), Var("num_kv_indices"), Const("num_heads", abbrev="h"),
          Const("head_dim_ckv", abbrev="ckv"), Const("head_dim_kpe", abbrev="kpe"), Const("page_size", abbrev="ps")],
    inputs=[Tensor("q_nope", ("total_q", "num_heads", "head_dim_ckv")), Tensor("q_pe", ("total_q", "num_heads", "head_dim_kpe")),
            Tensor("ckv_cache", ("num_pages", "page_size", "head_dim_ckv")), Tensor("kpe_cache", ("num_pages", "page_size", "head_dim_kpe")),
            Tensor("qo_indptr", ("len_indptr",), "int32", param="self._qo_host"), Tensor("kv_indptr", ("len_indptr",), "int32", param="self._kvp_host"),
            Tensor("kv_indices", ("num_kv_indices",), "int32", param="self._kv_indices"),
            Tensor("kv_len_arr", ("batch_size",), "int32", param="self._kvl_host"), Scalar("causal", "bool", param="self._causal"),
            Scalar("sm_scale", param="self._sm_scale")],
    outputs=[Tensor("output", ("total_q", "num_heads", "head_dim_ckv"), dtype_from="q_nope"), Tensor("lse", ("total_q", "num_heads"), dtype="float32")],
    reference=_mla_paged_reference, init=_mla_paged_init, tags=("attention", "mla", "paged"), constraints=("len_indptr == batch_size + 1",),
    description="Multi-head latent attention over paged compressed KV (matrix-absorbed form)", tolerance="bf16",
    test_sizes={"num_heads": 4, "head_dim_ckv": 512, "head_dim_kpe": 64, "page_size": 4, "batch_size": 3})


# ------------------------------------------------------------------ block-table entry points (TRT-LLM / cuDNN / XQA call styles)
def _block_table_decode_reference(query, k_cache, v_cache, block_tables, seq_lens, bmm1_scale=1.0, bmm2_scale=1.0):
    """HND pages: k_cache / v_cache [pages, Hkv, page_size, D]; request b reads pages block_tables[b, :ceil(len / page_size)]."""
    b, h, d = query.shape
    hkv, page_size = k_cache.shape[1], k_cache.shape[2]
    group = h // hkv
    out = torch.zeros(b, h, d, dtype=torch.float32, device=query.device)
    for i in range(b):
        n = int(seq_lens[i])
        pages = block_tables[i, : (n + page_size - 1) // page_size].long()
        k = k_cache[pages].permute(0, 2, 1, 3).reshape(-1, hkv, d)[:n].to(torch.float32).repeat_interleave(group, dim=1)
        v = v_cache[pages].permute(0, 2, 1, 3).reshape(-1, hkv, d)[:n].to(torch.float32).repeat_interleave(group, dim=1)
        logits = torch.einsum("hd,lhd->hl", query[i].to(torch.float32), k) * bmm1_scale
        out[i] = torch.einsum("hl,lhd->hd", torch.softmax(logits, -1), v) * bmm2_scale
    return out.to(query.dtype)


def _block_table_inputs(batch_size, num_qo_heads, num_kv_heads, head_dim, page_size, kv_len, device, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    lens = [int(x) for x in torch.randint(1, 6 * page_size, (batch_size,), generator=g)] if kv_len is None else [int(kv_len)] * batch_size
    per = [(n + page_size - 1) // page_size for n in lens]
    total = sum(per)
    ids = torch.randperm(total + 2, generator=g)[:total].int()
    table = torch.zeros(batch_size, max(per), dtype=torch.int32)
    o = 0
    for i, p in enumerate(per):
        table[i, :p] = ids[o:o + p]
        o += p
    mk = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16).to(device)  # noqa: E731
    return (mk(batch_size, num_qo_heads, head_dim), mk(total + 2, num_kv_heads, page_size, head_dim), mk(total + 2, num_kv_heads, page_size, head_dim),
            table.to(device), torch.tensor(lens, dtype=torch.int32, device=device), max(lens))


def _trtllm_decode_init(*, batch_size=8, num_qo_heads=32, num_kv_heads=8, head_dim=128, page_size=16, kv_len=None, device="cuda", seed=0):
    q, k, v, table, lens, mx = _block_table_inputs(batch_size, num_qo_heads, num_kv_heads, head_dim, page_size, kv_len, device, seed)
    return {"query": q, "kv_cache": (k, v), "workspace_buffer": torch.empty(32 << 20, dtype=torch.uint8, device=device), "block_tables": table,
            "seq_lens": lens, "max_seq_len": mx, "bmm1_scale": 1.0 / math.sqrt(head_dim), "bmm2_scale": 1.0, "kv_layout": "HND"}


_HND = ("num_pages", "num_kv_heads", "page_size", "head_dim")
trtllm_batch_decode_trace = TraceTemplate(
    op_type="gqa_paged", name_fmt="trtllm_batch_decode_h{num_qo_heads}_kv{num_kv_heads}_d{head_dim}_ps{page_size}",
    axes=[Var("batch_size"), Var("num_pages"), Var("max_pages_per_seq")] + _H + [Const("page_size", abbrev="ps")],
    inputs=[Tensor("query", ("batch_size", "num_qo_heads", "head_dim")), Tensor("k_cache", _HND, param="kv_cache", tuple_idx=0),
            Tensor("v_cache", _HND, param="kv_cache", tuple_idx=1), Tensor("block_tables", ("batch_size", "max_pages_per_seq"), "int32"),
            Tensor("seq_lens", ("batch_size",), "int32"), Scalar("bmm1_scale", description="softmax scale with q / k de-quantisation folded in"),
            Scalar("bmm2_scale", optional=True, description="output scale (v de-quantisation)")],
    outputs=[Tensor("output", ("batch_size", "num_qo_heads", "head_dim"), dtype_from="query")], reference=_block_table_decode_reference,
    init=_trtllm_decode_init, tags=("attention", "decode", "paged", "block_table"),
    description="Batched GQA decode addressed by a dense block table over HND pages (TRT-LLM call style)", tolerance="bf16", test_sizes=_SIZES)


def _cudnn_decode_reference(q, k_cache, v_cache, scale, actual_seq_lens_kv, block_tables):
    b, h, d = q.shape
    hkv, page_size = k_cache.shape[1], k_cache.shape[2]
    group = h // hkv
    lens = actual_seq_lens_kv.reshape(-1)
    out = torch.zeros(b, h, d, dtype=torch.float32, device=q.device)
    for i in range(b):
        n = int(lens[i])
        pages = block_tables[i, : (n + page_size - 1) // page_size].long()
        k = k_cache[pages].permute(0, 2, 1, 3).reshape(-1, hkv, d)[:n].to(torch.float32).repeat_interleave(group, dim=1)
        v = v_cache[pages].permute(0, 2, 1, 3).reshape(-1, hkv, d)[:n].to(torch.float32).repeat_interleave(group, dim=1)
        logits = torch.einsum("hd,lhd->hl", q[i].to(torch.float32), k) * scale
        out[i] = torch.einsum("hl,lhd->hd", torch.softmax(logits, -1), v)
    return out.to(q.dtype)


def _cudnn_decode_init(*, batch_size=8, num_qo_heads=32, num_kv_heads=8, head_dim=128, page_size=16, kv_len=None, device="cuda", seed=0):
    q, k, v, table, lens, mx = _block_table_inputs(batch_size, num_qo_heads, num_kv_heads, head_dim, page_size, kv_len, device, seed)
    return {"q": q, "k_cache": k, "v_cache": v, "scale": 1.0 / math.sqrt(head_dim), "workspace_buffer": torch.empty(32 << 20, dtype=torch.uint8, device=device),
            "max_sequence_kv": mx, "actual_seq_lens_kv": lens.view(-1, 1, 1, 1), "block_tables": table}


cudnn_batch_decode_trace = TraceTemplate(
    op_type="gqa_paged", name_fmt="cudnn_batch_decode_h{num_qo_heads}_kv{num_kv_heads}_d{head_dim}_ps{page_size}",
    axes=[Var("batch_size"), Var("num_pages"), Var("max_pages_per_seq")] + _H + [Const("page_size", abbrev="ps")],
    inputs=[Tensor("q", ("batch_size", "num_qo_heads", "head_dim")), Tensor("k_cache", _HND), Tensor("v_cache", _HND), Scalar("scale"),
            Tensor("actual_seq_lens_kv", ("batch_size", "one", "one_", "one__"), "int32"),
            Tensor("block_tables", ("batch_size", "max_pages_per_seq"), "int32")],
    outputs=[Tensor("output", ("batch_size", "num_qo_heads", "head_dim"), dtype_from="q")], reference=_cudnn_decode_reference,
    init=_cudnn_decode_init, tags=("attention", "decode", "paged", "block_table"), constraints=("one == 1", "one_ == 1", "one__ == 1"),
    description="cuDNN-style call signature of the paged decode kernel", tolerance="bf16", test_sizes=_SIZES)
