"""More attention templates (reference flashinfer/trace/templates/attention.py: POD, batch POD, multi-level cascade, variable
block-sparse, the TRT-LLM / cuDNN / fmha_v2 context entry points, dense MLA decode through the block-table call style, XQA MLA).

Unlike the reference's run()-templates (which approximate wrapper calls by one dense SDPA), the wrapper templates here read the
plan()-time tables off the wrapper object, so their references are exact and checked against the API by the generic test."""
import math

import torch

from ..template import Const, Scalar, TemplateDispatch, Tensor, TraceTemplate, Var
from .attention import (_gqa_paged_decode_reference, _gqa_paged_prefill_reference, _gqa_ragged_prefill_reference, _paged_tables,
                        _single_prefill_reference)
from .round2 import sparse_mla_decode_trace

_H = [Const("num_qo_heads", abbrev="h"), Const("num_kv_heads", abbrev="kv"), Const("head_dim", abbrev="d")]
_NHD = ("num_pages", "page_size", "num_kv_heads", "head_dim")
_HND = ("num_pages", "num_kv_heads", "page_size", "head_dim")
_SIZES = {"num_qo_heads": 4, "num_kv_heads": 2, "head_dim": 64, "page_size": 4}


def _mk(g, device):
    return lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16).to(device)


def _pair_compare(got, expected, kwargs):
    """(prefill output, decode output) pairs; either side may come back as (out, lse)."""
    for g_, e_ in zip(got, expected):
        g_ = g_[0] if isinstance(g_, (tuple, list)) else g_
        torch.testing.assert_close(g_.float(), e_.float(), atol=3e-2, rtol=3e-2)


# ------------------------------------------------------------------ POD: one ragged prefill request + a paged decode batch
def _pod_reference(q_p, k_p, v_p, q_d, k_cache_d, v_cache_d, kv_indptr_d, kv_indices_d, kv_last_page_len_d, causal_p):
    o_p = _single_prefill_reference(q_p, k_p, v_p, causal=bool(causal_p))
    o_d, _ = _gqa_paged_decode_reference(q_d, k_cache_d, v_cache_d, kv_indptr_d, kv_indices_d, kv_last_page_len_d)
    return o_p, o_d


def _pod_init(*, qo_len=1024, kv_len=1024, batch_size=32, num_qo_heads=32, num_kv_heads=8, head_dim=128, page_size=16, device="cuda", seed=0):
    import flashinfer_b200 as fi

    g = torch.Generator(device="cpu").manual_seed(seed)
    mk = _mk(g, device)
    lens = [int(x) for x in torch.randint(1, 6 * page_size, (batch_size,), generator=g)]
    indices, indptr, last, num_pages = _paged_tables(lens, page_size, device, g)
    w = fi.PODWithPagedKVCacheWrapper(torch.empty(32 << 20, dtype=torch.uint8, device=device), "NHD")
    w.plan(indptr, indices, last, num_qo_heads, num_kv_heads, head_dim, page_size, q_data_type=torch.bfloat16)
    return {"self": w, "q_p": mk(qo_len, num_qo_heads, head_dim), "k_p": mk(kv_len, num_kv_heads, head_dim), "v_p": mk(kv_len, num_kv_heads, head_dim),
            "q_d": mk(batch_size, num_qo_heads, head_dim),
            "paged_kv_cache_d": (mk(num_pages, page_size, num_kv_heads, head_dim), mk(num_pages, page_size, num_kv_heads, head_dim)), "causal_p": True}


pod_with_paged_kv_cache_run_trace = TraceTemplate(
    op_type="pod", name_fmt="pod_with_paged_kv_cache_h{num_qo_heads}_kv{num_kv_heads}_d{head_dim}_ps{page_size}",
    axes=[Var("qo_len"), Var("kv_len"), Var("batch_size"), Var("num_pages"), Var("len_indptr"), Var("num_kv_indices")] + _H + [Const("page_size", abbrev="ps")],
    inputs=[Tensor("q_p", ("qo_len", "num_qo_heads", "head_dim")), Tensor("k_p", ("kv_len", "num_kv_heads", "head_dim")),
            Tensor("v_p", ("kv_len", "num_kv_heads", "head_dim")), Tensor("q_d", ("batch_size", "num_qo_heads", "head_dim")),
            Tensor("k_cache_d", _NHD, param="paged_kv_cache_d", tuple_idx=0), Tensor("v_cache_d", _NHD, param="paged_kv_cache_d", tuple_idx=1),
            Tensor("kv_indptr_d", ("len_indptr",), "int32", param="self._dargs.0"), Tensor("kv_indices_d", ("num_kv_indices",), "int32", param="self._dargs.1"),
            Tensor("kv_last_page_len_d", ("batch_size",), "int32", param="self._dargs.2"), Scalar("causal_p", "bool")],
    outputs=[Tensor("output_p", ("qo_len", "num_qo_heads", "head_dim"), dtype_from="q_p"),
             Tensor("output_d", ("batch_size", "num_qo_heads", "head_dim"), dtype_from="q_d")],
    reference=_pod_reference, init=_pod_init, compare=_pair_compare, helpers=(_single_prefill_reference, _gqa_paged_decode_reference),
    tags=("attention", "pod", "prefill", "decode"), constraints=("len_indptr == batch_size + 1",),
    description="Prefill-on-decode: one ragged prefill request and a paged decode batch in one launch (SMs split by the cost model)",
    test_sizes=dict(_SIZES, qo_len=24, kv_len=40, batch_size=5))


# ------------------------------------------------------------------ batch POD: paged prefill batch + paged decode batch
def _batch_pod_reference(q_p, k_cache_p, v_cache_p, qo_indptr_p, kv_indptr_p, kv_indices_p, kv_last_page_len_p, causal, q_d, k_cache_d,
                         v_cache_d, kv_indptr_d, kv_indices_d, kv_last_page_len_d):
    o_p, _ = _gqa_paged_prefill_reference(q_p, k_cache_p, v_cache_p, qo_indptr_p, kv_indptr_p, kv_indices_p, kv_last_page_len_p, causal=bool(causal))
    o_d, _ = _gqa_paged_decode_reference(q_d, k_cache_d, v_cache_d, kv_indptr_d, kv_indices_d, kv_last_page_len_d)
    return o_p, o_d


def _batch_pod_init(*, prefill_batch=4, batch_size=32, num_qo_heads=32, num_kv_heads=8, head_dim=128, page_size=16, device="cuda", seed=0):
    import flashinfer_b200 as fi

    g = torch.Generator(device="cpu").manual_seed(seed)
    mk = _mk(g, device)
    q_lens = [int(x) for x in torch.randint(2, 3 * page_size, (prefill_batch,), generator=g)]
    kv_lens_p = [ql + int(x) for ql, x in zip(q_lens, torch.randint(0, 3 * page_size, (prefill_batch,), generator=g))]
    idx_p, indptr_p, last_p, pages_p = _paged_tables(kv_lens_p, page_size, device, g)
    lens_d = [int(x) for x in torch.randint(1, 6 * page_size, (batch_size,), generator=g)]
    idx_d, indptr_d, last_d, pages_d = _paged_tables(lens_d, page_size, device, g)
    qo_p = torch.tensor([0] + list(torch.tensor(q_lens).cumsum(0)), dtype=torch.int32, device=device)
    qo_d = torch.arange(batch_size + 1, dtype=torch.int32, device=device)
    w = fi.BatchPODWithPagedKVCacheWrapper(torch.empty(32 << 20, dtype=torch.uint8, device=device), "NHD")
    w.plan(qo_p, indptr_p, idx_p, last_p, qo_d, indptr_d, idx_d, last_d, num_qo_heads, num_kv_heads, head_dim, page_size,
           q_data_type=torch.bfloat16, causal_p=True)
    return {"self": w, "q_p": mk(sum(q_lens), num_qo_heads, head_dim),
            "paged_kv_cache_p": (mk(pages_p, page_size, num_kv_heads, head_dim), mk(pages_p, page_size, num_kv_heads, head_dim)),
            "q_d": mk(batch_size, num_qo_heads, head_dim),
            "paged_kv_cache_d": (mk(pages_d, page_size, num_kv_heads, head_dim), mk(pages_d, page_size, num_kv_heads, head_dim))}


_NHD_P = ("num_pages_p", "page_size", "num_kv_heads", "head_dim")
batch_pod_with_paged_kv_cache_run_trace = TraceTemplate(
    op_type="pod", name_fmt="batch_pod_with_paged_kv_cache_h{num_qo_heads}_kv{num_kv_heads}_d{head_dim}_ps{page_size}",
    axes=[Var("total_q_p"), Var("prefill_batch"), Var("len_indptr_p"), Var("num_kv_indices_p"), Var("num_pages_p"), Var("batch_size"), Var("num_pages"),
          Var("len_indptr"), Var("num_kv_indices")] + _H + [Const("page_size", abbrev="ps")],
    inputs=[Tensor("q_p", ("total_q_p", "num_qo_heads", "head_dim")), Tensor("k_cache_p", _NHD_P, param="paged_kv_cache_p", tuple_idx=0),
            Tensor("v_cache_p", _NHD_P, param="paged_kv_cache_p", tuple_idx=1),
            Tensor("qo_indptr_p", ("len_indptr_p",), "int32", param="self._prefill._qo_indptr_host"),
            Tensor("kv_indptr_p", ("len_indptr_p",), "int32", param="self._prefill._kv_indptr_host"),
            Tensor("kv_indices_p", ("num_kv_indices_p",), "int32", param="self._prefill._kv_indices"),
            Tensor("kv_last_page_len_p", ("prefill_batch",), "int32", param="self._prefill._kv_last_host"),
            Scalar("causal", "bool", param="self._prefill._causal"),
            Tensor("q_d", ("batch_size", "num_qo_heads", "head_dim")), Tensor("k_cache_d", _NHD, param="paged_kv_cache_d", tuple_idx=0),
            Tensor("v_cache_d", _NHD, param="paged_kv_cache_d", tuple_idx=1),
            Tensor("kv_indptr_d", ("len_indptr",), "int32", param="self._decode._kv_indptr_host"),
            Tensor("kv_indices_d", ("num_kv_indices",), "int32", param="self._decode._kv_indices"),
            Tensor("kv_last_page_len_d", ("batch_size",), "int32", param="self._decode._kv_last_host")],
    outputs=[Tensor("output_p", ("total_q_p", "num_qo_heads", "head_dim"), dtype_from="q_p"),
             Tensor("output_d", ("batch_size", "num_qo_heads", "head_dim"), dtype_from="q_d")],
    reference=_batch_pod_reference, init=_batch_pod_init, compare=_pair_compare, helpers=(_gqa_paged_prefill_reference, _gqa_paged_decode_reference),
    tags=("attention", "pod", "prefill", "decode", "paged"), constraints=("len_indptr == batch_size + 1", "len_indptr_p == prefill_batch + 1"),
    description="Batched prefill-on-decode: a paged prefill batch and a paged decode batch in one launch",
    test_sizes=dict(_SIZES, prefill_batch=3, batch_size=5))


# ------------------------------------------------------------------ two-level cascade (shared prefix + unique suffix)
def _cascade_reference(q, k_cache, v_cache, qo_indptr_0, kv_indptr_0, kv_indices_0, kv_last_page_len_0, qo_indptr_1, kv_indptr_1, kv_indices_1,
                       kv_last_page_len_1, causal, sm_scale):
    """Every query token attends the union of the level-0 (shared) KV of the group it belongs to and its own level-1 (unique) KV;
    ``causal`` applies to level 1 only (the shared prefix precedes every query token)."""
    t, h, d = q.shape
    page_size, hkv = k_cache.shape[1], k_cache.shape[2]
    grp = h // hkv
    out = torch.zeros(t, h, d, dtype=torch.float32, device=q.device)

    def level_keys(tok, qo_indptr, kv_indptr, kv_indices, last):
        i = int(torch.searchsorted(qo_indptr.long().cpu(), torch.tensor(tok), right=True)) - 1
        pages = kv_indices[int(kv_indptr[i]): int(kv_indptr[i + 1])].long()
        if pages.numel() == 0:
            return k_cache.new_zeros(0, hkv, d).float(), v_cache.new_zeros(0, hkv, d).float(), tok - int(qo_indptr[i]), int(qo_indptr[i + 1] - qo_indptr[i])
        n = (pages.numel() - 1) * page_size + int(last[i])
        return (k_cache[pages].reshape(-1, hkv, d)[:n].float(), v_cache[pages].reshape(-1, hkv, d)[:n].float(), tok - int(qo_indptr[i]),
                int(qo_indptr[i + 1] - qo_indptr[i]))

    for tok in range(t):
        k0, v0, _, _ = level_keys(tok, qo_indptr_0, kv_indptr_0, kv_indices_0, kv_last_page_len_0)
        k1, v1, j, qlen = level_keys(tok, qo_indptr_1, kv_indptr_1, kv_indices_1, kv_last_page_len_1)
        if causal:
            keep = k1.shape[0] - qlen + j + 1
            k1, v1 = k1[:keep], v1[:keep]
        k = torch.cat([k0, k1]).repeat_interleave(grp, dim=1)
        v = torch.cat([v0, v1]).repeat_interleave(grp, dim=1)
        p = torch.softmax(torch.einsum("hd,nhd->hn", q[tok].float(), k) * sm_scale, -1)
        out[tok] = torch.einsum("hn,nhd->hd", p, v)
    return out.to(q.dtype)


def _cascade_init(*, num_groups=2, requests_per_group=3, num_qo_heads=32, num_kv_heads=8, head_dim=128, page_size=16, device="cuda", seed=0):
    import flashinfer_b200 as fi

    g = torch.Generator(device="cpu").manual_seed(seed)
    mk = _mk(g, device)
    b = num_groups * requests_per_group
    q_lens = [int(x) for x in torch.randint(1, 6, (b,), generator=g)]
    shared = [int(x) for x in torch.randint(page_size, 4 * page_size, (num_groups,), generator=g)]
    unique = [ql + int(x) for ql, x in zip(q_lens, torch.randint(0, 3 * page_size, (b,), generator=g))]
    per = [(n + page_size - 1) // page_size for n in shared + unique]
    ids = torch.randperm(sum(per) + 2, generator=g)[: sum(per)].int()
    split = list(torch.tensor([0] + per).cumsum(0))
    chunks = [ids[int(split[i]): int(split[i + 1])] for i in range(len(per))]
    cum = lambda xs: torch.tensor([0] + list(torch.tensor(xs).cumsum(0)), dtype=torch.int32)  # noqa: E731
    qo1 = cum(q_lens)
    qo0 = qo1[::requests_per_group].clone()
    kv0, kv1 = cum(per[:num_groups]), cum(per[num_groups:])
    last = lambda ns: torch.tensor([(n - 1) % page_size + 1 for n in ns], dtype=torch.int32)  # noqa: E731
    w = fi.MultiLevelCascadeAttentionWrapper(2, torch.empty(32 << 20, dtype=torch.uint8, device=device), "NHD")
    w.plan([qo0.to(device), qo1.to(device)], [kv0.to(device), kv1.to(device)],
           [torch.cat(chunks[:num_groups]).to(device), torch.cat(chunks[num_groups:]).to(device)], [last(shared).to(device), last(unique).to(device)],
           num_qo_heads, num_kv_heads, head_dim, page_size, causal=True, q_data_type=torch.bfloat16)
    n_pages = sum(per) + 2
    return {"self": w, "q": mk(sum(q_lens), num_qo_heads, head_dim),
            "paged_kv_cache": (mk(n_pages, page_size, num_kv_heads, head_dim), mk(n_pages, page_size, num_kv_heads, head_dim))}


def _level(i):
    pre = f"self._batch_prefill_wrappers.{i}."
    return [Tensor(f"qo_indptr_{i}", (f"len_indptr_{i}",), "int32", param=pre + "_qo_indptr_host"),
            Tensor(f"kv_indptr_{i}", (f"len_indptr_{i}",), "int32", param=pre + "_kv_indptr_host"),
            Tensor(f"kv_indices_{i}", (f"num_kv_indices_{i}",), "int32", param=pre + "_kv_indices"),
            Tensor(f"kv_last_page_len_{i}", (f"batch_{i}",), "int32", param=pre + "_kv_last_host")]


multi_level_cascade_run_trace = TraceTemplate(
    op_type="cascade_attention", name_fmt="multi_level_cascade_h{num_qo_heads}_kv{num_kv_heads}_d{head_dim}_ps{page_size}",
    axes=[Var("total_q"), Var("num_pages"), Var("len_indptr_0"), Var("num_kv_indices_0"), Var("batch_0"), Var("len_indptr_1"), Var("num_kv_indices_1"),
          Var("batch_1")] + _H + [Const("page_size", abbrev="ps")],
    inputs=[Tensor("q", ("total_q", "num_qo_heads", "head_dim")), Tensor("k_cache", _NHD, param="paged_kv_cache", tuple_idx=0),
            Tensor("v_cache", _NHD, param="paged_kv_cache", tuple_idx=1)] + _level(0) + _level(1)
    + [Scalar("causal", "bool", param="self._batch_prefill_wrappers.1._causal"), Scalar("sm_scale", param="self._batch_prefill_wrappers.1._sm_scale")],
    outputs=[Tensor("output", ("total_q", "num_qo_heads", "head_dim"), dtype_from="q")], reference=_cascade_reference, init=_cascade_init,
    tags=("attention", "cascade", "shared-prefix", "paged"), constraints=("len_indptr_0 == batch_0 + 1", "len_indptr_1 == batch_1 + 1"),
    description="Two-level cascade attention: per-level paged prefill passes merged by log-sum-exp (shared prefix level + unique suffix level)",
    tolerance="bf16", test_sizes=_SIZES)


# ------------------------------------------------------------------ variable block-sparse attention
def _var_block_sparse_reference(q, k, v, block_mask_map, block_row_sz, block_col_sz, sm_scale):
    """q [Hq, Sq, D], k / v [Hkv, Skv, D]; for kv head h, row block i attends the column blocks j with block_mask_map[h, i, j];
    block sizes differ per kv head (block_row_sz [Hkv, MB], block_col_sz [Hkv, NB])."""
    hq, sq, d = q.shape
    hkv = k.shape[0]
    grp = hq // hkv
    out = torch.zeros(hq, sq, d, dtype=torch.float32, device=q.device)
    for h in range(hkv):
        rows = torch.cat([torch.zeros(1, dtype=torch.long), block_row_sz[h].long().cumsum(0)])
        cols = torch.cat([torch.zeros(1, dtype=torch.long), block_col_sz[h].long().cumsum(0)])
        for i in range(block_mask_map.shape[1]):
            sel = [torch.arange(int(cols[j]), int(cols[j + 1])) for j in range(block_mask_map.shape[2]) if bool(block_mask_map[h, i, j])]
            if not sel or rows[i] == rows[i + 1]:
                continue
            idx = torch.cat(sel).to(q.device)
            qs = q[h * grp:(h + 1) * grp, int(rows[i]): int(rows[i + 1])].float()
            p = torch.softmax(torch.einsum("gqd,nd->gqn", qs, k[h, idx].float()) * sm_scale, -1)
            out[h * grp:(h + 1) * grp, int(rows[i]): int(rows[i + 1])] = torch.einsum("gqn,nd->gqd", p, v[h, idx].float())
    return out.to(q.dtype)


def _var_block_sparse_init(*, num_row_blocks=5, num_col_blocks=6, num_qo_heads=8, num_kv_heads=2, head_dim=128, seq_q=96, seq_kv=160,
                           device="cuda", seed=0):
    import flashinfer_b200 as fi

    g = torch.Generator(device="cpu").manual_seed(seed)
    mk = _mk(g, device)

    def sizes(total, parts):
        cuts = torch.sort(torch.randperm(total - 1, generator=g)[: parts - 1] + 1).values
        edges = torch.cat([torch.zeros(1, dtype=torch.long), cuts, torch.tensor([total])])
        return (edges[1:] - edges[:-1]).int()

    rs = torch.stack([sizes(seq_q, num_row_blocks) for _ in range(num_kv_heads)])
    cs = torch.stack([sizes(seq_kv, num_col_blocks) for _ in range(num_kv_heads)])
    mask = torch.rand(num_kv_heads, num_row_blocks, num_col_blocks, generator=g) < 0.5
    mask[..., 0] = True                                   # every row block sees at least one column block
    w = fi.VariableBlockSparseAttentionWrapper(torch.empty(32 << 20, dtype=torch.uint8, device=device))
    w.plan(mask.to(device), rs.to(device), cs.to(device), num_qo_heads, num_kv_heads, head_dim, q_data_type=torch.bfloat16)
    return {"self": w, "q": mk(num_qo_heads, seq_q, head_dim), "k": mk(num_kv_heads, seq_kv, head_dim), "v": mk(num_kv_heads, seq_kv, head_dim)}


variable_block_sparse_attention_run_trace = TraceTemplate(
    op_type="block_sparse", name_fmt="variable_block_sparse_attention_h{num_qo_heads}_kv{num_kv_heads}_d{head_dim}",
    axes=[Var("seq_q"), Var("seq_kv"), Var("num_row_blocks"), Var("num_col_blocks")] + _H,
    inputs=[Tensor("q", ("num_qo_heads", "seq_q", "head_dim")), Tensor("k", ("num_kv_heads", "seq_kv", "head_dim")),
            Tensor("v", ("num_kv_heads", "seq_kv", "head_dim")),
            Tensor("block_mask_map", ("num_kv_heads", "num_row_blocks", "num_col_blocks"), "bool", param="self._block_mask_map"),
            Tensor("block_row_sz", ("num_kv_heads", "num_row_blocks"), "int64", param="self._block_row_sz"),
            Tensor("block_col_sz", ("num_kv_heads", "num_col_blocks"), "int64", param="self._block_col_sz"),
            Scalar("sm_scale", param="self._sm_scale")],
    outputs=[Tensor("output", ("num_qo_heads", "seq_q", "head_dim"), dtype_from="q")], reference=_var_block_sparse_reference,
    init=_var_block_sparse_init, tags=("attention", "sparse", "variable-block"), tolerance="bf16",
    description="Block-sparse attention with per-head variable block sizes (column blocks expanded to a token-granular page list)",
    test_sizes={"num_qo_heads": 4, "num_kv_heads": 2, "head_dim": 64, "seq_q": 40, "seq_kv": 56})


# ------------------------------------------------------------------ block-table context (prefill) entry points
def _block_table_context_reference(query, k_cache, v_cache, block_tables, seq_lens, cum_seq_lens_q, bmm1_scale, bmm2_scale):
    """HND pages [pages, Hkv, page_size, D]; request b owns query rows cum_seq_lens_q[b]:cum_seq_lens_q[b+1] - the LAST tokens of its
    seq_lens[b] cached tokens (causal)."""
    h, d = query.shape[1:]
    hkv, page_size = k_cache.shape[1], k_cache.shape[2]
    grp = h // hkv
    out = torch.zeros(query.shape, dtype=torch.float32, device=query.device)
    for b in range(seq_lens.numel()):
        qs, qe, n = int(cum_seq_lens_q[b]), int(cum_seq_lens_q[b + 1]), int(seq_lens[b])
        pages = block_tables[b, : (n + page_size - 1) // page_size].long()
        k = k_cache[pages].permute(0, 2, 1, 3).reshape(-1, hkv, d)[:n].float().repeat_interleave(grp, dim=1)
        v = v_cache[pages].permute(0, 2, 1, 3).reshape(-1, hkv, d)[:n].float().repeat_interleave(grp, dim=1)
        logits = torch.einsum("qhd,nhd->hqn", query[qs:qe].float(), k) * bmm1_scale
        qpos = torch.arange(qe - qs, device=query.device)[:, None] + (n - (qe - qs))
        logits = logits.masked_fill(torch.arange(n, device=query.device)[None, :] > qpos, float("-inf"))
        out[qs:qe] = torch.einsum("hqn,nhd->qhd", torch.softmax(logits, -1), v) * bmm2_scale
    return out.to(query.dtype)


def _block_table_context_inputs(batch_size, num_qo_heads, num_kv_heads, head_dim, page_size, device, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    mk = _mk(g, device)
    q_lens = [int(x) for x in torch.randint(1, 3 * page_size, (batch_size,), generator=g)]
    kv_lens = [ql + int(x) for ql, x in zip(q_lens, torch.randint(0, 4 * page_size, (batch_size,), generator=g))]
    per = [(n + page_size - 1) // page_size for n in kv_lens]
    ids = torch.randperm(sum(per) + 2, generator=g)[: sum(per)].int()
    table = torch.zeros(batch_size, max(per), dtype=torch.int32)
    o = 0
    for i, p in enumerate(per):
        table[i, :p] = ids[o:o + p]
        o += p
    cum = lambda xs: torch.tensor([0] + list(torch.tensor(xs).cumsum(0)), dtype=torch.int32, device=device)  # noqa: E731
    n_pages = sum(per) + 2
    return {"q": mk(sum(q_lens), num_qo_heads, head_dim), "k": mk(n_pages, num_kv_heads, page_size, head_dim),
            "v": mk(n_pages, num_kv_heads, page_size, head_dim), "table": table.to(device), "q_lens": q_lens, "kv_lens": kv_lens,
            "cum_q": cum(q_lens), "cum_kv": cum(kv_lens)}


def _trtllm_context_init(*, batch_size=4, num_qo_heads=32, num_kv_heads=8, head_dim=128, page_size=16, device="cuda", seed=0):
    t = _block_table_context_inputs(batch_size, num_qo_heads, num_kv_heads, head_dim, page_size, device, seed)
    return {"query": t["q"], "kv_cache": (t["k"], t["v"]), "workspace_buffer": torch.empty(32 << 20, dtype=torch.uint8, device=device),
            "block_tables": t["table"], "seq_lens": torch.tensor(t["kv_lens"], dtype=torch.int32, device=device), "max_q_len": max(t["q_lens"]),
            "max_kv_len": max(t["kv_lens"]), "bmm1_scale": 1.0 / math.sqrt(head_dim), "bmm2_scale": 1.0, "batch_size": batch_size,
            "cum_seq_lens_q": t["cum_q"], "cum_seq_lens_kv": t["cum_kv"], "kv_layout": "HND"}


def _first_output(got, expected, kwargs):
    out = got[0][0] if isinstance(got[0], (tuple, list)) else got[0]
    torch.testing.assert_close(out.float(), expected[0].float(), atol=3e-2, rtol=3e-2)


trtllm_batch_context_trace = TraceTemplate(
    op_type="gqa_paged", name_fmt="trtllm_batch_context_h{num_qo_heads}_kv{num_kv_heads}_d{head_dim}_ps{page_size}",
    axes=[Var("total_q"), Var("batch_size"), Var("num_pages"), Var("max_pages_per_seq"), Var("len_cum")] + _H + [Const("page_size", abbrev="ps")],
    inputs=[Tensor("query", ("total_q", "num_qo_heads", "head_dim")), Tensor("k_cache", _HND, param="kv_cache", tuple_idx=0),
            Tensor("v_cache", _HND, param="kv_cache", tuple_idx=1), Tensor("block_tables", ("batch_size", "max_pages_per_seq"), "int32"),
            Tensor("seq_lens", ("batch_size",), "int32"), Tensor("cum_seq_lens_q", ("len_cum",), "int32"), Scalar("bmm1_scale"), Scalar("bmm2_scale")],
    outputs=[Tensor("output", ("total_q", "num_qo_heads", "head_dim"), dtype_from="query")], reference=_block_table_context_reference,
    init=_trtllm_context_init, compare=_first_output, tags=("attention", "prefill", "paged", "block_table"), constraints=("len_cum == batch_size + 1",),
    description="Causal context (prefill / append) attention addressed by a dense block table over HND pages (TRT-LLM call style)",
    test_sizes=_SIZES)


def _cudnn_prefill_reference(q, k_cache, v_cache, scale, actual_seq_lens_q, actual_seq_lens_kv, block_tables, causal):
    """cuDNN call style: packed q rows in request order, lengths as [B, 1, 1, 1] tensors, HND pages + block table."""
    lq = actual_seq_lens_q.reshape(-1).long()
    cum_q = torch.cat([torch.zeros(1, dtype=torch.long, device=lq.device), lq.cumsum(0)])
    h, d = q.shape[1:]
    hkv, page_size = k_cache.shape[1], k_cache.shape[2]
    grp = h // hkv
    out = torch.zeros(q.shape, dtype=torch.float32, device=q.device)
    for b in range(lq.numel()):
        qs, qe, n = int(cum_q[b]), int(cum_q[b + 1]), int(actual_seq_lens_kv.reshape(-1)[b])
        pages = block_tables[b, : (n + page_size - 1) // page_size].long()
        k = k_cache[pages].permute(0, 2, 1, 3).reshape(-1, hkv, d)[:n].float().repeat_interleave(grp, dim=1)
        v = v_cache[pages].permute(0, 2, 1, 3).reshape(-1, hkv, d)[:n].float().repeat_interleave(grp, dim=1)
        logits = torch.einsum("qhd,nhd->hqn", q[qs:qe].float(), k) * scale
        if causal:
            qpos = torch.arange(qe - qs, device=q.device)[:, None] + (n - (qe - qs))
            logits = logits.masked_fill(torch.arange(n, device=q.device)[None, :] > qpos, float("-inf"))
        out[qs:qe] = torch.einsum("hqn,nhd->qhd", torch.softmax(logits, -1), v)
    return out.to(q.dtype)


def _cudnn_prefill_init(*, batch_size=4, num_qo_heads=32, num_kv_heads=8, head_dim=128, page_size=16, device="cuda", seed=0):
    t = _block_table_context_inputs(batch_size, num_qo_heads, num_kv_heads, head_dim, page_size, device, seed)
    lens = lambda xs: torch.tensor(xs, dtype=torch.int32, device=device).view(-1, 1, 1, 1)  # noqa: E731
    return {"q": t["q"], "k_cache": t["k"], "v_cache": t["v"], "scale": 1.0 / math.sqrt(head_dim),
            "workspace_buffer": torch.empty(32 << 20, dtype=torch.uint8, device=device), "max_token_per_sequence": max(t["q_lens"]),
            "max_sequence_kv": max(t["kv_lens"]), "actual_seq_lens_q": lens(t["q_lens"]), "actual_seq_lens_kv": lens(t["kv_lens"]),
            "block_tables": t["table"], "causal": True}


cudnn_batch_prefill_trace = TraceTemplate(
    op_type="gqa_paged", name_fmt="cudnn_batch_prefill_h{num_qo_heads}_kv{num_kv_heads}_d{head_dim}_ps{page_size}",
    axes=[Var("total_q"), Var("batch_size"), Var("num_pages"), Var("max_pages_per_seq")] + _H + [Const("page_size", abbrev="ps")],
    inputs=[Tensor("q", ("total_q", "num_qo_heads", "head_dim")), Tensor("k_cache", _HND), Tensor("v_cache", _HND), Scalar("scale"),
            Tensor("actual_seq_lens_q", ("batch_size", "one", "one", "one"), "int32"), Tensor("actual_seq_lens_kv", ("batch_size", "one", "one", "one"), "int32"),
            Tensor("block_tables", ("batch_size", "max_pages_per_seq"), "int32"), Scalar("causal", "bool")],
    outputs=[Tensor("output", ("total_q", "num_qo_heads", "head_dim"), dtype_from="q")], reference=_cudnn_prefill_reference,
    init=_cudnn_prefill_init, compare=_first_output, tags=("attention", "prefill", "paged", "cudnn-style"), constraints=("one == 1",),
    description="Batched prefill in the cuDNN call style (4-d length tensors, HND pages + block table), served by the tcgen05 prefill kernel",
    test_sizes=_SIZES)


# ------------------------------------------------------------------ DeepSeek context attention (192 / 128 head dims), ragged
def _ragged_deepseek_reference(query, key, value, cum_seq_lens_q, cum_seq_lens_kv, bmm1_scale, bmm2_scale, is_causal):
    out, lse = _gqa_ragged_prefill_reference(query, key, value, cum_seq_lens_q, cum_seq_lens_kv, causal=bool(is_causal), sm_scale=bmm1_scale)
    return (out.float() * bmm2_scale).to(query.dtype)


def _ragged_deepseek_init(*, batch_size=4, num_heads=128, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    mk = _mk(g, device)
    lens = [int(x) for x in torch.randint(1, 40, (batch_size,), generator=g)]
    cum = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=device)
    n = sum(lens)
    return {"query": mk(n, num_heads, 192), "key": mk(n, num_heads, 192), "value": mk(n, num_heads, 128),
            "workspace_buffer": torch.empty(32 << 20, dtype=torch.uint8, device=device), "seq_lens": torch.tensor(lens, dtype=torch.int32, device=device),
            "max_q_len": max(lens), "max_kv_len": max(lens), "bmm1_scale": 1.0 / math.sqrt(192.0), "bmm2_scale": 1.0, "o_sf_scale": -1.0,
            "batch_size": batch_size, "window_left": -1, "cum_seq_lens_q": cum, "cum_seq_lens_kv": cum, "is_causal": True}


trtllm_ragged_attention_deepseek_trace = TraceTemplate(
    op_type="gqa_ragged", name_fmt="trtllm_ragged_attention_deepseek_h{num_heads}_qk{head_dim_qk}_vo{head_dim_vo}",
    axes=[Var("total_q"), Var("total_kv"), Var("len_cum"), Const("num_heads", abbrev="h"), Const("head_dim_qk", abbrev="qk"), Const("head_dim_vo", abbrev="vo")],
    inputs=[Tensor("query", ("total_q", "num_heads", "head_dim_qk")), Tensor("key", ("total_kv", "num_heads", "head_dim_qk")),
            Tensor("value", ("total_kv", "num_heads", "head_dim_vo")), Tensor("cum_seq_lens_q", ("len_cum",), "int32"),
            Tensor("cum_seq_lens_kv", ("len_cum",), "int32"), Scalar("bmm1_scale"), Scalar("bmm2_scale"), Scalar("is_causal", "bool")],
    outputs=[Tensor("output", ("total_q", "num_heads", "head_dim_vo"), dtype_from="query")], reference=_ragged_deepseek_reference,
    init=_ragged_deepseek_init, compare=_first_output, helpers=(_gqa_ragged_prefill_reference,), tags=("attention", "prefill", "ragged", "deepseek"),
    description="DeepSeek MLA context attention in its expanded form: ragged q / k with 192-wide heads, v with 128-wide heads",
    test_sizes={"num_heads": 4})


def _fmha_v2_deepseek_reference(query, key, value, scale_softmax):
    """query / key [B, S, H, 192], value [B, S, H, 128], causal inside every sequence; scale_softmax 0 means 1 / sqrt(192)."""
    b, s, h, dqk = query.shape
    sm = scale_softmax if scale_softmax else dqk ** -0.5
    logits = torch.einsum("bqhd,bkhd->bhqk", query.float(), key.float()) * sm
    logits = logits.masked_fill(torch.ones(s, s, dtype=torch.bool, device=query.device).triu(1), float("-inf"))
    return torch.einsum("bhqk,bkhd->bqhd", torch.softmax(logits, -1), value.float()).to(query.dtype)


def _fmha_v2_deepseek_init(*, batch_size=2, seq_len=256, num_heads=128, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    mk = _mk(g, device)
    return {"query": mk(batch_size, seq_len, num_heads, 192), "key": mk(batch_size, seq_len, num_heads, 192), "value": mk(batch_size, seq_len, num_heads, 128),
            "out": torch.empty(batch_size, seq_len, num_heads, 128, dtype=torch.bfloat16, device=device), "num_heads": num_heads, "head_dim": 192,
            "seq_len": seq_len, "scale_softmax": 0.0}


fmha_v2_prefill_deepseek_trace = TraceTemplate(
    op_type="gqa_ragged", name_fmt="fmha_v2_prefill_deepseek_h{num_heads}_qk{head_dim_qk}_vo{head_dim_vo}",
    axes=[Var("batch_size"), Var("seq_len"), Const("num_heads", abbrev="h"), Const("head_dim_qk", abbrev="qk"), Const("head_dim_vo", abbrev="vo")],
    inputs=[Tensor("query", ("batch_size", "seq_len", "num_heads", "head_dim_qk")), Tensor("key", ("batch_size", "seq_len", "num_heads", "head_dim_qk")),
            Tensor("value", ("batch_size", "seq_len", "num_heads", "head_dim_vo")), Scalar("scale_softmax")],
    outputs=[Tensor("out", ("batch_size", "seq_len", "num_heads", "head_dim_vo"), dtype_from="query", param="out")],
    reference=_fmha_v2_deepseek_reference, init=_fmha_v2_deepseek_init, tags=("attention", "prefill", "deepseek", "fixed-length"), tolerance="bf16",
    description="DeepSeek-R1 context attention over equal-length sequences (an sm_120 fmha_v2 kernel in the reference; the tcgen05 192 / 128 kernel here)",
    test_sizes={"num_heads": 4, "seq_len": 24})


# ------------------------------------------------------------------ dense MLA decode through the block-table call style (+ sparse dispatch)
def _mla_block_table_reference(query, kv_cache, block_tables, seq_lens, bmm1_scale, bmm2_scale):
    """query [B, q_len, H, 576] = [nope-absorbed 512 | rope 64]; kv_cache [pages, 1, page_size, 576]; the q_len query tokens of a request
    are its LAST tokens (causal among themselves: speculative / multi-token decode)."""
    b, ql, h, _ = query.shape
    kv = kv_cache.squeeze(1) if kv_cache.dim() == 4 else kv_cache
    page_size = kv.shape[1]
    out = torch.zeros(b, ql, h, 512, dtype=torch.float32, device=query.device)
    for i in range(b):
        n = int(seq_lens[i])
        rows = kv[block_tables[i, : (n + page_size - 1) // page_size].long()].reshape(-1, 576)[:n].float()
        logits = torch.einsum("qhd,nd->qhn", query[i].float(), rows) * bmm1_scale
        qpos = torch.arange(ql, device=query.device)[:, None] + (n - ql)
        logits = logits.masked_fill((torch.arange(n, device=query.device)[None, :] > qpos)[:, None, :], float("-inf"))
        out[i] = torch.einsum("qhn,nd->qhd", torch.softmax(logits, -1), rows[:, :512]) * bmm2_scale
    return out.to(query.dtype)


def _mla_block_table_inputs(batch_size, q_len, num_heads, page_size, device, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    lens = [q_len + int(x) for x in torch.randint(0, 5 * page_size, (batch_size,), generator=g)]
    per = [(n + page_size - 1) // page_size for n in lens]
    ids = torch.randperm(sum(per) + 2, generator=g)[: sum(per)].int()
    table = torch.zeros(batch_size, max(per), dtype=torch.int32)
    o = 0
    for i, p in enumerate(per):
        table[i, :p] = ids[o:o + p]
        o += p
    kv = torch.randn(sum(per) + 2, 1, page_size, 576, generator=g).clamp(-1, 1).to(torch.bfloat16)
    q = (torch.randn(batch_size, q_len, num_heads, 576, generator=g) * 0.5).to(torch.bfloat16)
    return q.to(device), kv.to(device), table.to(device), torch.tensor(lens, dtype=torch.int32, device=device), max(lens)


def _mla_block_table_init(*, batch_size=4, q_len=1, num_heads=128, page_size=64, device="cuda", seed=0):
    q, kv, table, lens, mx = _mla_block_table_inputs(batch_size, q_len, num_heads, page_size, device, seed)
    return {"query": q, "kv_cache": kv, "workspace_buffer": torch.zeros(32 << 20, dtype=torch.uint8, device=device), "qk_nope_head_dim": 128,
            "kv_lora_rank": 512, "qk_rope_head_dim": 64, "block_tables": table, "seq_lens": lens, "max_seq_len": mx,
            "bmm1_scale": 1.0 / math.sqrt(192.0), "bmm2_scale": 1.0}


trtllm_batch_decode_mla_trace = TraceTemplate(
    op_type="mla_paged", name_fmt="trtllm_batch_decode_mla_h{num_heads}_ps{page_size}",
    axes=[Var("batch_size"), Var("q_len"), Var("num_pages"), Var("max_pages_per_seq"), Const("num_heads", abbrev="h"), Const("page_size", abbrev="ps")],
    inputs=[Tensor("query", ("batch_size", "q_len", "num_heads", "qk_dim")), Tensor("kv_cache", ("num_pages", "one", "page_size", "qk_dim")),
            Tensor("block_tables", ("batch_size", "max_pages_per_seq"), "int32"), Tensor("seq_lens", ("batch_size",), "int32"), Scalar("bmm1_scale"),
            Scalar("bmm2_scale")],
    outputs=[Tensor("out", ("batch_size", "q_len", "num_heads", "kv_lora_rank"), dtype_from="query")], reference=_mla_block_table_reference,
    init=_mla_block_table_init, tags=("mla", "decode", "paged", "block_table"), tolerance="bf16",
    constraints=("one == 1", "qk_dim == 576", "kv_lora_rank == 512"),
    description="Absorbed MLA decode (512-d latent + 64-d rope key in one 576-wide cache row) addressed by a dense block table",
    test_sizes={"num_heads": 4, "page_size": 16, "q_len": 2})

trtllm_batch_decode_mla_trace_dispatch = TemplateDispatch(
    [trtllm_batch_decode_mla_trace, sparse_mla_decode_trace],
    lambda bound: sparse_mla_decode_trace if bound.get("sparse_mla_top_k") else trtllm_batch_decode_mla_trace)


def _xqa_mla_reference(q, k_cache, page_table, seq_lens, q_scale, kv_scale=None):
    """q [B, 1, H, 576]; k_cache [pages, page_size, 576] (the latent cache is both K and V); softmax scale q_scale * kv_scale / sqrt(192)."""
    b, _, h, _ = q.shape
    page_size = k_cache.shape[-2]
    kv = k_cache.reshape(-1, page_size, 576)
    sm = q_scale * (kv_scale if kv_scale is not None else 1.0) / math.sqrt(192.0)
    out = torch.zeros(b, 1, h, 512, dtype=torch.float32, device=q.device)
    for i in range(b):
        n = int(seq_lens.reshape(-1)[i])
        rows = kv[page_table[i, : (n + page_size - 1) // page_size].long()].reshape(-1, 576)[:n].float()
        p = torch.softmax(q[i, 0].float() @ rows.t() * sm, -1)
        out[i, 0] = p @ rows[:, :512]
    return out.to(q.dtype)


def _xqa_mla_init(*, batch_size=4, num_heads=128, page_size=64, device="cuda", seed=0):
    q, kv, table, lens, _ = _mla_block_table_inputs(batch_size, 1, num_heads, page_size, device, seed)
    return {"q": q, "k_cache": kv.squeeze(1), "v_cache": kv.squeeze(1)[..., :512], "page_table": table, "seq_lens": lens.view(-1, 1),
            "output": torch.empty(batch_size, 1, num_heads, 512, dtype=torch.bfloat16, device=device),
            "workspace_buffer": torch.zeros(32 << 20, dtype=torch.uint8, device=device), "q_scale": 1.0, "kv_scale": None}


xqa_mla_trace = TraceTemplate(
    op_type="mla_paged", name_fmt="xqa_mla_h{num_heads}_ps{page_size}",
    axes=[Var("batch_size"), Var("num_pages"), Var("max_pages"), Const("num_heads", abbrev="h"), Const("page_size", abbrev="ps")],
    inputs=[Tensor("q", ("batch_size", "one", "num_heads", "qk_dim")), Tensor("k_cache", ("num_pages", "page_size", "qk_dim")),
            Tensor("page_table", ("batch_size", "max_pages"), "int32"), Tensor("seq_lens", ("batch_size", "one"), "int32"), Scalar("q_scale"),
            Scalar("kv_scale", optional=True)],
    outputs=[Tensor("out", ("batch_size", "one", "num_heads", "kv_lora_rank"), dtype_from="q", param="output")], reference=_xqa_mla_reference,
    init=_xqa_mla_init, tags=("mla", "decode", "paged", "xqa"), tolerance="bf16", constraints=("one == 1", "qk_dim == 576", "kv_lora_rank == 512"),
    description="XQA-style MLA decode: per-request page table rows + sequence lengths over a 576-wide latent cache (tcgen05 MLA kernel)",
    test_sizes={"num_heads": 4, "page_size": 16})


# ------------------------------------------------------------------ XQA batch decode entry points (block table; NHD pages by default)
def _xqa_batch_decode_reference(query, k_cache, v_cache, block_tables, seq_lens, bmm1_scale, bmm2_scale):
    """NHD pages [pages, page_size, Hkv, D] (the XQA entry point's default layout); one query token per request."""
    b, h, d = query.shape
    page_size, hkv = k_cache.shape[1], k_cache.shape[2]
    grp = h // hkv
    out = torch.zeros(b, h, d, dtype=torch.float32, device=query.device)
    for i in range(b):
        n = int(seq_lens[i])
        pages = block_tables[i, : (n + page_size - 1) // page_size].long()
        k = k_cache[pages].reshape(-1, hkv, d)[:n].float().repeat_interleave(grp, dim=1)
        v = v_cache[pages].reshape(-1, hkv, d)[:n].float().repeat_interleave(grp, dim=1)
        p = torch.softmax(torch.einsum("hd,nhd->hn", query[i].float(), k) * bmm1_scale, -1)
        out[i] = torch.einsum("hn,nhd->hd", p, v) * bmm2_scale
    return out.to(query.dtype)


def _xqa_batch_decode_init(*, batch_size=8, num_qo_heads=32, num_kv_heads=8, head_dim=128, page_size=16, device="cuda", seed=0):
    t = _block_table_context_inputs(batch_size, num_qo_heads, num_kv_heads, head_dim, page_size, device, seed)
    g = torch.Generator(device="cpu").manual_seed(seed + 3)
    return {"query": _mk(g, device)(batch_size, num_qo_heads, head_dim), "kv_cache": (t["k"].transpose(1, 2).contiguous(), t["v"].transpose(1, 2).contiguous()),
            "workspace_buffer": torch.zeros(32 << 20, dtype=torch.uint8, device=device), "block_tables": t["table"],
            "seq_lens": torch.tensor(t["kv_lens"], dtype=torch.int32, device=device), "max_seq_len": max(t["kv_lens"]),
            "bmm1_scale": 1.0 / math.sqrt(head_dim), "bmm2_scale": 1.0}


xqa_batch_decode_trace = TraceTemplate(
    op_type="gqa_paged", name_fmt="xqa_batch_decode_h{num_qo_heads}_kv{num_kv_heads}_d{head_dim}_ps{page_size}",
    axes=[Var("batch_size"), Var("num_pages"), Var("max_pages_per_seq")] + _H + [Const("page_size", abbrev="ps")],
    inputs=[Tensor("query", ("batch_size", "num_qo_heads", "head_dim")), Tensor("k_cache", _NHD, param="kv_cache", tuple_idx=0),
            Tensor("v_cache", _NHD, param="kv_cache", tuple_idx=1), Tensor("block_tables", ("batch_size", "max_pages_per_seq"), "int32"),
            Tensor("seq_lens", ("batch_size",), "int32"), Scalar("bmm1_scale"), Scalar("bmm2_scale")],
    outputs=[Tensor("output", ("batch_size", "num_qo_heads", "head_dim"), dtype_from="query")], reference=_xqa_batch_decode_reference,
    init=_xqa_batch_decode_init, tags=("attention", "decode", "paged", "block_table", "xqa"), tolerance="bf16",
    description="XQA-named batch decode entry point (NHD pages by default); the tcgen05 paged decode kernel on B200", test_sizes=_SIZES)


def _xqa_batch_mla_init(*, batch_size=4, q_len=1, num_heads=128, page_size=64, device="cuda", seed=0):
    kw = _mla_block_table_init(batch_size=batch_size, q_len=q_len, num_heads=num_heads, page_size=page_size, device=device, seed=seed)
    return kw


xqa_batch_decode_mla_trace = TraceTemplate(
    op_type="mla_paged", name_fmt="xqa_batch_decode_mla_h{num_heads}_ps{page_size}",
    axes=[Var("batch_size"), Var("q_len"), Var("num_pages"), Var("max_pages_per_seq"), Const("num_heads", abbrev="h"), Const("page_size", abbrev="ps")],
    inputs=[Tensor("query", ("batch_size", "q_len", "num_heads", "qk_dim")), Tensor("kv_cache", ("num_pages", "one", "page_size", "qk_dim")),
            Tensor("block_tables", ("batch_size", "max_pages_per_seq"), "int32"), Tensor("seq_lens", ("batch_size",), "int32"), Scalar("bmm1_scale"),
            Scalar("bmm2_scale")],
    outputs=[Tensor("out", ("batch_size", "q_len", "num_heads", "kv_lora_rank"), dtype_from="query")], reference=_mla_block_table_reference,
    init=_xqa_batch_mla_init, tags=("mla", "decode", "paged", "block_table", "xqa"), tolerance="bf16",
    constraints=("one == 1", "qk_dim == 576", "kv_lora_rank == 512"),
    description="XQA-named MLA batch decode (an sm_120 kernel in the reference; the tcgen05 MLA kernel here)",
    test_sizes={"num_heads": 4, "page_size": 16, "q_len": 1})

__all__ = [n for n in dir() if n.endswith("_trace") or n.endswith("_trace_dispatch")]
