"""Paged-KV maintenance templates (reference flashinfer/trace/templates/page.py).  Memory movement only: tolerance exact."""
import torch

from ..template import Const, Scalar, Tensor, TraceTemplate, Var


def _append_paged_kv_cache_reference(append_key, append_value, batch_indices, positions, k_cache, v_cache, kv_indices, kv_indptr,
                                     kv_last_page_len=None, kv_layout="NHD"):
    """Token t of request b = batch_indices[t] lands in page kv_indices[kv_indptr[b] + positions[t] // page_size], slot
    positions[t] % page_size.  NHD caches are [pages, page_size, heads, dim]; HND caches [pages, heads, page_size, dim]."""
    k_out, v_out = k_cache.clone(), v_cache.clone()
    page_size = k_cache.shape[1] if kv_layout == "NHD" else k_cache.shape[2]
    for t in range(append_key.shape[0]):
        b, pos = int(batch_indices[t]), int(positions[t])
        page = int(kv_indices[int(kv_indptr[b]) + pos // page_size])
        slot = pos % page_size
        if kv_layout == "NHD":
            k_out[page, slot], v_out[page, slot] = append_key[t], append_value[t]
        else:
            k_out[page, :, slot], v_out[page, :, slot] = append_key[t], append_value[t]
    return k_out, v_out


def _paged_layout(batch, page_size, lens, device, g):
    pages_per = [(n + page_size - 1) // page_size for n in lens]
    total = sum(pages_per)
    perm = torch.randperm(total + 3, generator=g)[:total].int()        # scattered, not identity, page ids
    indptr = torch.tensor([0] + list(torch.tensor(pages_per).cumsum(0)), dtype=torch.int32)
    last = torch.tensor([(n - 1) % page_size + 1 for n in lens], dtype=torch.int32)
    return perm.to(device), indptr.to(device), last.to(device), total + 3


def _append_init(*, nnz=24, num_kv_heads=8, head_dim=128, page_size=16, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    lens_new = [nnz // 3, nnz - nnz // 3]                              # tokens appended per request
    lens_old = [5, 0]                                                  # tokens already cached
    lens = [a + b for a, b in zip(lens_old, lens_new)]
    kv_indices, kv_indptr, last, num_pages = _paged_layout(2, page_size, lens, device, g)
    bi = torch.cat([torch.full((n,), i, dtype=torch.int32) for i, n in enumerate(lens_new)]).to(device)
    pos = torch.cat([torch.arange(o, o + n, dtype=torch.int32) for o, n in zip(lens_old, lens_new)]).to(device)
    mk = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16).to(device)  # noqa: E731
    return {"append_key": mk(nnz, num_kv_heads, head_dim), "append_value": mk(nnz, num_kv_heads, head_dim), "batch_indices": bi,
            "positions": pos, "paged_kv_cache": (mk(num_pages, page_size, num_kv_heads, head_dim), mk(num_pages, page_size, num_kv_heads, head_dim)),
            "kv_indices": kv_indices, "kv_indptr": kv_indptr, "kv_last_page_len": last, "kv_layout": "NHD"}


_CACHE = ("num_pages", "page_size", "num_kv_heads", "head_dim")
append_paged_kv_cache_trace = TraceTemplate(
    op_type="page", name_fmt="append_paged_kv_cache_h{num_kv_heads}_d{head_dim}_ps{page_size}",
    axes=[Var("nnz"), Var("batch_size"), Var("num_pages"), Var("len_indptr"), Var("num_kv_indices"), Const("num_kv_heads", abbrev="h"),
          Const("head_dim", abbrev="d"), Const("page_size", abbrev="ps")],
    inputs=[Tensor("append_key", ("nnz", "num_kv_heads", "head_dim")), Tensor("append_value", ("nnz", "num_kv_heads", "head_dim")),
            Tensor("batch_indices", ("nnz",), "int32"), Tensor("positions", ("nnz",), "int32"),
            Tensor("k_cache", _CACHE, param="paged_kv_cache", tuple_idx=0), Tensor("v_cache", _CACHE, param="paged_kv_cache", tuple_idx=1),
            Tensor("kv_indices", ("num_kv_indices",), "int32"), Tensor("kv_indptr", ("len_indptr",), "int32"),
            Tensor("kv_last_page_len", ("batch_size",), "int32"), Scalar("kv_layout", "str", optional=True)],
    outputs=[Tensor("k_cache_out", _CACHE, dtype_from="append_key", param="paged_kv_cache", tuple_idx=0),
             Tensor("v_cache_out", _CACHE, dtype_from="append_key", param="paged_kv_cache", tuple_idx=1)],
    reference=_append_paged_kv_cache_reference, init=_append_init, tags=("page", "inplace"),
    constraints=("len_indptr == batch_size + 1",), description="Scatter new K/V tokens into their pages (NHD layout shown)",
    tolerance="exact", test_sizes={"num_kv_heads": 2, "head_dim": 64, "page_size": 4})


def _append_paged_mla_kv_cache_reference(append_ckv, append_kpe, batch_indices, positions, ckv_cache, kpe_cache, kv_indices, kv_indptr):
    c_out, p_out = ckv_cache.clone(), kpe_cache.clone()
    page_size = ckv_cache.shape[1]
    for t in range(append_ckv.shape[0]):
        b, pos = int(batch_indices[t]), int(positions[t])
        page = int(kv_indices[int(kv_indptr[b]) + pos // page_size])
        c_out[page, pos % page_size], p_out[page, pos % page_size] = append_ckv[t], append_kpe[t]
    return c_out, p_out


def _append_mla_init(*, nnz=24, head_dim_ckv=512, head_dim_kpe=64, page_size=16, device="cuda", seed=0):
    kw = _append_init(nnz=nnz, num_kv_heads=1, head_dim=8, page_size=page_size, device=device, seed=seed)
    g = torch.Generator(device="cpu").manual_seed(seed + 3)
    num_pages = kw["paged_kv_cache"][0].shape[0]
    mk = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16).to(device)  # noqa: E731
    return {"append_ckv": mk(nnz, head_dim_ckv), "append_kpe": mk(nnz, head_dim_kpe), "batch_indices": kw["batch_indices"],
            "positions": kw["positions"], "ckv_cache": mk(num_pages, page_size, head_dim_ckv), "kpe_cache": mk(num_pages, page_size, head_dim_kpe),
            "kv_indices": kw["kv_indices"], "kv_indptr": kw["kv_indptr"], "kv_last_page_len": kw["kv_last_page_len"]}


append_paged_mla_kv_cache_trace = TraceTemplate(
    op_type="page", name_fmt="append_paged_mla_kv_cache_ckv{head_dim_ckv}_kpe{head_dim_kpe}_ps{page_size}",
    axes=[Var("nnz"), Var("batch_size"), Var("num_pages"), Var("len_indptr"), Var("num_kv_indices"), Const("head_dim_ckv", abbrev="ckv"),
          Const("head_dim_kpe", abbrev="kpe"), Const("page_size", abbrev="ps")],
    inputs=[Tensor("append_ckv", ("nnz", "head_dim_ckv")), Tensor("append_kpe", ("nnz", "head_dim_kpe")),
            Tensor("batch_indices", ("nnz",), "int32"), Tensor("positions", ("nnz",), "int32"),
            Tensor("ckv_cache", ("num_pages", "page_size", "head_dim_ckv")), Tensor("kpe_cache", ("num_pages", "page_size", "head_dim_kpe")),
            Tensor("kv_indices", ("num_kv_indices",), "int32"), Tensor("kv_indptr", ("len_indptr",), "int32"),
            Tensor("kv_last_page_len", ("batch_size",), "int32", optional=True)],
    outputs=[Tensor("ckv_cache_out", ("num_pages", "page_size", "head_dim_ckv"), dtype_from="append_ckv", param="ckv_cache"),
             Tensor("kpe_cache_out", ("num_pages", "page_size", "head_dim_kpe"), dtype_from="append_kpe", param="kpe_cache")],
    reference=_append_paged_mla_kv_cache_reference, init=_append_mla_init, tags=("page", "mla", "inplace"),
    constraints=("len_indptr == batch_size + 1",), description="Scatter compressed-KV and rope-key tokens into the MLA page pools",
    tolerance="exact", test_sizes={"head_dim_ckv": 64, "head_dim_kpe": 16, "page_size": 4})


def _get_batch_indices_positions_reference(append_indptr, seq_lens, nnz):
    """Token t of request b (append_indptr[b] <= t < append_indptr[b+1]) sits at position
    seq_lens[b] - (append_indptr[b+1] - append_indptr[b]) + (t - append_indptr[b])."""
    bi = torch.zeros(nnz, dtype=torch.int32, device=append_indptr.device)
    pos = torch.zeros(nnz, dtype=torch.int32, device=append_indptr.device)
    for b in range(append_indptr.numel() - 1):
        s, e = int(append_indptr[b]), int(append_indptr[b + 1])
        bi[s:e] = b
        pos[s:e] = int(seq_lens[b]) - (e - s) + torch.arange(e - s, dtype=torch.int32, device=append_indptr.device)
    return bi, pos


def _bip_init(*, batch_size=4, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    new = torch.randint(1, 9, (batch_size,), generator=g)
    indptr = torch.cat([torch.zeros(1, dtype=torch.long), new.cumsum(0)]).int()
    return {"append_indptr": indptr.to(device), "seq_lens": (new + torch.randint(0, 50, (batch_size,), generator=g)).int().to(device),
            "nnz": int(indptr[-1])}


get_batch_indices_positions_trace = TraceTemplate(
    op_type="page", name_fmt="get_batch_indices_positions", axes=[Var("batch_size"), Var("len_indptr"), Var("nnz")],
    inputs=[Tensor("append_indptr", ("len_indptr",), "int32"), Tensor("seq_lens", ("batch_size",), "int32"), Scalar("nnz", "int32")],
    outputs=[Tensor("batch_indices", ("nnz",), dtype="int32"), Tensor("positions", ("nnz",), dtype="int32")],
    reference=_get_batch_indices_positions_reference, init=_bip_init, tags=("page",), constraints=("len_indptr == batch_size + 1",),
    description="Per-token (request id, absolute position) for an append of ragged new tokens", tolerance="exact")
