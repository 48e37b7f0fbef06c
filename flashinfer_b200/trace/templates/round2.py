"""Templates of the ops added in round 2: the flagship decode linear and its epilogues, the cluster top-k family and its index
transforms, SM-constrained GEMM, greedy argmax sampling, the sparse-MLA row gather path and the legacy MoE all-to-all index
preparation (reference counterparts: flashinfer/trace/templates/{gemm,sampling,attention,moe}.py)."""
import math

import torch

from ..template import Const, Scalar, Tensor, TraceTemplate, Var


# ------------------------------------------------------------------ decode linear (csrc/gemm/decode_linear_sm100.cu)
def _decode_linear_plain_reference(x, w, row_sumsq, eps):
    """out = rsqrt(row_sumsq / K + eps)[:, None] * (x @ w^T): RMSNorm folded into the GEMM (the gain lives in w)."""
    acc = x.to(torch.float32) @ w.to(torch.float32).t()
    rstd = torch.rsqrt(row_sumsq[: x.shape[0]].to(torch.float32) / x.shape[1] + eps)
    return (acc * rstd[:, None]).to(x.dtype)


def _decode_linear_plain_init(*, num_tokens=64, out_features=6144, in_features=4096, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(num_tokens, in_features, generator=g).to(torch.bfloat16)
    return {"x": x.to(device), "w": (torch.randn(out_features, in_features, generator=g) / in_features ** 0.5).to(torch.bfloat16).to(device),
            "row_sumsq": x.float().pow(2).sum(-1).to(device), "eps": 1e-5}


decode_linear_trace = TraceTemplate(
    op_type="gemm", name_fmt="decode_linear_n{out_features}_k{in_features}",
    axes=[Var("num_tokens"), Const("out_features", abbrev="n"), Const("in_features", abbrev="k")],
    inputs=[Tensor("x", ("num_tokens", "in_features")), Tensor("w", ("out_features", "in_features")),
            Tensor("row_sumsq", ("num_tokens",), optional=True), Scalar("eps", "float32", optional=True)],
    outputs=[Tensor("out", ("num_tokens", "out_features"), dtype_from="x")], reference=_decode_linear_plain_reference,
    init=_decode_linear_plain_init, tags=("gemm", "decode", "small-m", "rmsnorm-folded"), tolerance="cos",
    description="Small-M (<= 64 tokens) weight-streaming GEMM of a decode step with the RMSNorm folded in (plain epilogue)",
    test_sizes={"out_features": 96, "in_features": 128})


# ------------------------------------------------------------------ top-k family (csrc/elementwise/topk.cu)
def _topk_clusters_reference(logits, top_k):
    return torch.topk(logits.to(torch.float32), top_k, dim=-1, sorted=False).indices.to(torch.int32)


def _topk_clusters_init(*, num_rows=4, row_len=131072, top_k=2048, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return {"logits": torch.randn(num_rows, row_len, generator=g).to(device), "top_k": top_k}


def _topk_set_compare(got, expected, kwargs):
    (idx,), (ref,) = got, expected
    x = kwargs["logits"].float()
    assert idx.shape == ref.shape
    assert torch.equal(x.gather(1, idx.long()).sort(-1).values, x.gather(1, ref.long()).sort(-1).values), "selected values differ from top-k"
    assert (idx.long().sort(-1).values.diff(dim=-1) > 0).all(), "an index was selected twice"


topk_clusters_exact_trace = TraceTemplate(
    op_type="topk", name_fmt="topk_clusters_exact_n{row_len}_k{top_k}", axes=[Var("num_rows"), Const("row_len", abbrev="n"), Const("top_k", abbrev="k")],
    inputs=[Tensor("logits", ("num_rows", "row_len")), Scalar("top_k", "int32")],
    outputs=[Tensor("indices", ("num_rows", "top_k"), dtype="int32")], reference=_topk_clusters_reference, init=_topk_clusters_init,
    compare=_topk_set_compare, tags=("topk", "cluster", "dsa"),
    description="Exact top-k indices per row, one row per thread-block cluster (DSA indexer: few long rows)", test_sizes={"row_len": 777, "top_k": 33})


def _page_table_transform_reference(input, src_page_table, lengths, k):
    """out[i, j] = src_page_table[i, idx_j] for the top-k idx of input[i, :lengths[i]] (index order, -1 padded)."""
    rows = input.shape[0]
    out = torch.full((rows, k), -1, dtype=torch.int32, device=input.device)
    for r in range(rows):
        n = int(lengths[r])
        kk = min(k, n)
        sel = torch.topk(input[r, :n].to(torch.float32), kk).indices.sort().values
        out[r, :kk] = src_page_table[r, sel].to(torch.int32)
    return out


def _page_table_transform_init(*, num_rows=4, row_len=4096, k=256, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    lengths = torch.randint(max(1, k // 2), row_len + 1, (num_rows,), generator=g).to(torch.int32)
    lengths[0] = row_len
    return {"input": torch.randn(num_rows, row_len, generator=g).to(device),
            "src_page_table": torch.randint(0, 1 << 20, (num_rows, row_len), generator=g).to(torch.int32).to(device),
            "lengths": lengths.to(device), "k": k}


def _sorted_rows_compare(got, expected, kwargs):
    (g,), (e,) = got, expected
    assert g.shape == e.shape and torch.equal(g.sort(-1).values, e.sort(-1).values)


top_k_page_table_transform_trace = TraceTemplate(
    op_type="topk", name_fmt="top_k_page_table_transform_n{row_len}_k{k}", axes=[Var("num_rows"), Const("row_len", abbrev="n"), Const("k")],
    inputs=[Tensor("input", ("num_rows", "row_len")), Tensor("src_page_table", ("num_rows", "row_len"), dtype="int32"),
            Tensor("lengths", ("num_rows",), dtype="int32"), Scalar("k", "int32")],
    outputs=[Tensor("out", ("num_rows", "k"), dtype="int32")], reference=_page_table_transform_reference, init=_page_table_transform_init,
    compare=_sorted_rows_compare, tags=("topk", "dsa", "page-table"),
    description="Top-k over the valid prefix of every row fused with the page-table lookup of the selected positions", test_sizes={"row_len": 300, "k": 16})


def _ragged_transform_reference(input, offsets, lengths, k):
    rows = input.shape[0]
    out = torch.full((rows, k), -1, dtype=torch.int32, device=input.device)
    for r in range(rows):
        n = int(lengths[r])
        kk = min(k, n)
        sel = torch.topk(input[r, :n].to(torch.float32), kk).indices.sort().values
        out[r, :kk] = (sel + int(offsets[r])).to(torch.int32)
    return out


def _ragged_transform_init(*, num_rows=4, row_len=4096, k=256, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    lengths = torch.randint(max(1, k // 2), row_len + 1, (num_rows,), generator=g).to(torch.int32)
    return {"input": torch.randn(num_rows, row_len, generator=g).to(device),
            "offsets": (torch.arange(num_rows) * row_len).to(torch.int32).to(device), "lengths": lengths.to(device), "k": k}


top_k_ragged_transform_trace = TraceTemplate(
    op_type="topk", name_fmt="top_k_ragged_transform_n{row_len}_k{k}", axes=[Var("num_rows"), Const("row_len", abbrev="n"), Const("k")],
    inputs=[Tensor("input", ("num_rows", "row_len")), Tensor("offsets", ("num_rows",), dtype="int32"),
            Tensor("lengths", ("num_rows",), dtype="int32"), Scalar("k", "int32")],
    outputs=[Tensor("out", ("num_rows", "k"), dtype="int32")], reference=_ragged_transform_reference, init=_ragged_transform_init,
    compare=_sorted_rows_compare, tags=("topk", "dsa", "ragged"),
    description="Top-k over the valid prefix of every row, emitted as indices into the ragged (concatenated) KV", test_sizes={"row_len": 300, "k": 16})


# ------------------------------------------------------------------ SM-constrained GEMM
def _gemm_persistent_reference(a, b):
    return (a.to(torch.float32) @ b.to(torch.float32)).to(a.dtype)


def _gemm_persistent_init(*, m=4096, n=4096, k=1024, num_sms=32, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    w = (torch.randn(n, k, generator=g) / k ** 0.5).to(torch.bfloat16).to(device)
    return {"a": (torch.randn(m, k, generator=g) * 0.5).to(torch.bfloat16).to(device), "b": w.t(), "num_sms": num_sms}


gemm_persistent_trace = TraceTemplate(
    op_type="gemm", name_fmt="sm_constraint_gemm_n{n}_k{k}", axes=[Var("m"), Const("n"), Const("k")],
    inputs=[Tensor("a", ("m", "k")), Tensor("b", ("k", "n")), Scalar("num_sms", "int32", optional=True)],
    outputs=[Tensor("c", ("m", "n"), dtype_from="a")], reference=_gemm_persistent_reference, init=_gemm_persistent_init, tolerance="cos",
    tags=("gemm", "persistent", "sm-constraint"), description="a @ b on a persistent grid of at most num_sms CTAs", test_sizes={"n": 96, "k": 64})


# ------------------------------------------------------------------ greedy sampling
def _local_argmax_reference(logits):
    return logits.to(torch.float32).argmax(-1)


def _local_argmax_init(*, batch_size=64, vocab_size=128256, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return {"logits": torch.randn(batch_size, vocab_size, generator=g).to(torch.bfloat16).to(device)}


def _argmax_compare(got, expected, kwargs):
    x = kwargs["logits"].float()
    assert torch.equal(x.gather(1, got[0][:, None].long()), x.gather(1, expected[0][:, None].long())), "argmax value differs"


local_argmax_trace = TraceTemplate(
    op_type="sampling", name_fmt="argmax_v{vocab_size}", axes=[Var("batch_size"), Const("vocab_size", abbrev="v")],
    inputs=[Tensor("logits", ("batch_size", "vocab_size"))], outputs=[Tensor("tokens", ("batch_size",), dtype="int64")],
    reference=_local_argmax_reference, init=_local_argmax_init, compare=_argmax_compare, tags=("sampling", "greedy"),
    description="Greedy sampling: row-wise argmax (lowest index on ties)", test_sizes={"vocab_size": 1000})


# ------------------------------------------------------------------ legacy MoE all-to-all: local gather
def _moe_local_gather_reference(recv_rank_cum_sum, local_gather_indices, gathered_expert_ids, gathered_scales, local_expert_ids, local_scales,
                                expert_count):
    n = int(recv_rank_cum_sum[-1])
    ids = torch.full_like(local_expert_ids, expert_count)
    sc = torch.zeros_like(local_scales)
    ids[:n] = gathered_expert_ids[local_gather_indices[:n].long()]
    sc[:n] = gathered_scales[local_gather_indices[:n].long()]
    return ids, sc


def _moe_local_gather_init(*, max_token_count_per_rank=64, ep_size=4, top_k=8, expert_count=64, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    alloc = max_token_count_per_rank * ep_size
    cum = torch.randint(0, max_token_count_per_rank + 1, (ep_size,), generator=g).cumsum(0).to(torch.int32)
    return {"recv_rank_cum_sum": cum.to(device), "local_gather_indices": torch.randint(0, alloc, (alloc,), generator=g).to(torch.int32).to(device),
            "gathered_expert_ids": torch.randint(0, expert_count, (alloc, top_k), generator=g).to(torch.int32).to(device),
            "gathered_scales": torch.rand(alloc, top_k, generator=g).to(device),
            "local_expert_ids": torch.empty(alloc, top_k, dtype=torch.int32, device=device), "local_scales": torch.empty(alloc, top_k, device=device),
            "max_token_count_per_rank": max_token_count_per_rank, "expert_count": expert_count, "top_k": top_k, "ep_rank": 1, "ep_size": ep_size}


moe_local_gather_trace = TraceTemplate(
    op_type="moe_comm", name_fmt="moe_local_gather_ep{ep_size}_k{top_k}", axes=[Var("alloc_tokens"), Const("ep_size", abbrev="ep"), Const("top_k", abbrev="k")],
    inputs=[Tensor("recv_rank_cum_sum", ("ep_size",), dtype="int32"), Tensor("local_gather_indices", ("alloc_tokens",), dtype="int32"),
            Tensor("gathered_expert_ids", ("alloc_tokens", "top_k"), dtype="int32"), Tensor("gathered_scales", ("alloc_tokens", "top_k")),
            Tensor("local_expert_ids", ("alloc_tokens", "top_k"), dtype="int32"), Tensor("local_scales", ("alloc_tokens", "top_k")),
            Scalar("expert_count", "int32")],
    outputs=[Tensor("local_expert_ids_out", ("alloc_tokens", "top_k"), dtype="int32", param="local_expert_ids"),
             Tensor("local_scales_out", ("alloc_tokens", "top_k"), dtype_from="gathered_scales", param="local_scales")],
    reference=_moe_local_gather_reference, init=_moe_local_gather_init, tolerance="exact", tags=("moe", "all-to-all", "legacy", "inplace"),
    description="Legacy MoE all-to-all: routing rows of the received tokens gathered from the all-gathered tables (invalid rows = expert_count / 0)",
    test_sizes={"ep_size": 4, "top_k": 3})


# ------------------------------------------------------------------ sparse MLA decode
def _sparse_mla_reference(query, kv_cache, block_tables, bmm1_scale):
    b, ql, h, _ = query.shape
    flat = kv_cache.reshape(-1, kv_cache.shape[-1]).to(torch.float32)
    out = torch.zeros(b, ql, h, 512, dtype=torch.float32, device=query.device)
    for i in range(b):
        for j in range(ql):
            sel = block_tables[i, j][block_tables[i, j] >= 0].long()
            rows = flat[sel]
            p = torch.softmax(query[i, j].to(torch.float32) @ rows.t() * bmm1_scale, -1)
            out[i, j] = p @ rows[:, :512]
    return out.to(query.dtype)


def _sparse_mla_init(*, batch_size=2, q_len=1, num_heads=128, top_k=2048, seq_len=4096, page_size=64, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    n_pages = batch_size * ((seq_len + page_size - 1) // page_size)
    kv = torch.randn(n_pages, page_size, 576, generator=g).clamp(-1, 1).to(torch.bfloat16)
    q = (torch.randn(batch_size, q_len, num_heads, 576, generator=g) * 0.5).to(torch.bfloat16)
    idx = torch.stack([torch.stack([torch.randperm(seq_len, generator=g)[:top_k] + b * ((seq_len + page_size - 1) // page_size) * page_size
                                    for _ in range(q_len)]) for b in range(batch_size)]).to(torch.int32)
    return {"query": q.to(device), "kv_cache": kv.unsqueeze(1).to(device), "workspace_buffer": torch.zeros(8 << 20, dtype=torch.uint8, device=device),
            "qk_nope_head_dim": 128, "kv_lora_rank": 512, "qk_rope_head_dim": 64, "block_tables": idx.to(device),
            "seq_lens": torch.full((batch_size,), seq_len, dtype=torch.int32, device=device), "max_seq_len": seq_len, "sparse_mla_top_k": top_k,
            "bmm1_scale": 1.0 / math.sqrt(192.0)}


sparse_mla_decode_trace = TraceTemplate(
    op_type="mla", name_fmt="sparse_mla_decode_h{num_heads}_topk{top_k}", axes=[Var("batch_size"), Var("q_len"), Const("num_heads", abbrev="h"), Const("top_k", abbrev="topk"),
                                                                              Var("num_pages"), Var("page_size")],
    inputs=[Tensor("query", ("batch_size", "q_len", "num_heads", "qk_dim")), Tensor("kv_cache", ("num_pages", "one", "page_size", "qk_dim")),
            Tensor("block_tables", ("batch_size", "q_len", "top_k"), dtype="int32"), Scalar("bmm1_scale", "float32")],
    outputs=[Tensor("out", ("batch_size", "q_len", "num_heads", "kv_lora_rank"), dtype_from="query")],
    reference=_sparse_mla_reference, init=_sparse_mla_init, tags=("mla", "decode", "sparse", "dsa"), tolerance="bf16",
    constraints=("one == 1", "qk_dim == 576", "kv_lora_rank == 512"),
    description="Sparse (top-k) MLA decode: every query token attends the KV rows listed in block_tables (-1 = unused)",
    test_sizes={"num_heads": 4, "top_k": 48, "seq_len": 100, "page_size": 32})

__all__ = [n for n in dir() if n.endswith("_trace")]


# ------------------------------------------------------------------ block-sparse attention (BSR mask, wrapper state from plan())
def _block_sparse_reference(q, k, v, bsr_indptr, bsr_indices, R, C, sm_scale):
    """Dense oracle: row block i attends the column blocks bsr_indices[bsr_indptr[i]:bsr_indptr[i + 1]] (no intra-block mask)."""
    M, hq, d = q.shape
    N, hkv = k.shape[0], k.shape[1]
    mask = torch.zeros(M, N, dtype=torch.bool, device=q.device)
    for i in range(M // R):
        for j in bsr_indices[int(bsr_indptr[i]):int(bsr_indptr[i + 1])].tolist():
            mask[i * R:(i + 1) * R, j * C:(j + 1) * C] = True
    g = hq // hkv
    kk = k.to(torch.float32).repeat_interleave(g, dim=1)
    vv = v.to(torch.float32).repeat_interleave(g, dim=1)
    logits = torch.einsum("mhd,nhd->hmn", q.to(torch.float32), kk) * sm_scale
    logits = logits.masked_fill(~mask[None], float("-inf"))
    return torch.einsum("hmn,nhd->mhd", torch.softmax(logits, -1), vv).to(q.dtype)


def _block_sparse_init(*, num_row_blocks=6, num_col_blocks=7, R=16, C=16, num_qo_heads=8, num_kv_heads=2, head_dim=128, device="cuda", seed=0):
    import flashinfer_b200 as fi

    g = torch.Generator(device="cpu").manual_seed(seed)
    dense = torch.rand(num_row_blocks, num_col_blocks, generator=g) < 0.4
    dense[:, 0] = True
    indptr = torch.zeros(num_row_blocks + 1, dtype=torch.int32)
    indptr[1:] = dense.sum(1).cumsum(0)
    indices = dense.nonzero()[:, 1].int()
    M, N = num_row_blocks * R, num_col_blocks * C
    mk = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16).to(device)  # noqa: E731
    w = fi.BlockSparseAttentionWrapper(torch.empty(32 << 20, dtype=torch.uint8, device=device))
    w.plan(indptr, indices, M, N, R, C, num_qo_heads, num_kv_heads, head_dim, q_data_type=torch.bfloat16)
    return {"self": w, "q": mk(M, num_qo_heads, head_dim), "k": mk(N, num_kv_heads, head_dim), "v": mk(N, num_kv_heads, head_dim)}


block_sparse_attention_trace = TraceTemplate(
    op_type="block_sparse", name_fmt="block_sparse_h{num_qo_heads}_kv{num_kv_heads}_d{head_dim}",
    axes=[Var("M"), Var("N"), Var("len_indptr"), Var("nnz_blocks"), Const("num_qo_heads", abbrev="h"), Const("num_kv_heads", abbrev="kv"),
          Const("head_dim", abbrev="d")],
    inputs=[Tensor("q", ("M", "num_qo_heads", "head_dim")), Tensor("k", ("N", "num_kv_heads", "head_dim")),
            Tensor("v", ("N", "num_kv_heads", "head_dim")), Tensor("bsr_indptr", ("len_indptr",), "int32", param="self._bsr_indptr"),
            Tensor("bsr_indices", ("nnz_blocks",), "int32", param="self._bsr_indices"), Scalar("R", "int32", param="self._R"),
            Scalar("C", "int32", param="self._C"), Scalar("sm_scale", param="self._sm_scale")],
    outputs=[Tensor("output", ("M", "num_qo_heads", "head_dim"), dtype_from="q")], reference=_block_sparse_reference, init=_block_sparse_init,
    tags=("attention", "block-sparse"), tolerance="bf16",
    description="Attention under a fixed-size block-sparse (BSR) mask: every row block is a request whose KV pages are its non-zero column blocks",
    test_sizes={"R": 4, "C": 8, "num_qo_heads": 4, "num_kv_heads": 2, "head_dim": 32})


# ------------------------------------------------------------------ BatchAttention (mixed prefill / decode batch)
def _batch_attention_reference(q, k_cache, v_cache, qo_indptr, kv_indptr, kv_indices, kv_len_arr, causal, sm_scale):
    h, d = q.shape[1:]
    page_size, hkv = k_cache.shape[1], k_cache.shape[2]
    g = h // hkv
    out = torch.zeros(q.shape, dtype=torch.float32, device=q.device)
    lse = torch.full((q.shape[0], h), float("-inf"), dtype=torch.float32, device=q.device)
    for b in range(qo_indptr.numel() - 1):
        qs, qe = int(qo_indptr[b]), int(qo_indptr[b + 1])
        n = int(kv_len_arr[b])
        if qe == qs or n == 0:
            continue
        pages = kv_indices[int(kv_indptr[b]): int(kv_indptr[b + 1])].long()
        k = k_cache[pages].reshape(-1, hkv, d)[:n].to(torch.float32).repeat_interleave(g, dim=1)
        v = v_cache[pages].reshape(-1, hkv, d)[:n].to(torch.float32).repeat_interleave(g, dim=1)
        logits = torch.einsum("qhd,nhd->hqn", q[qs:qe].to(torch.float32), k) * sm_scale
        if causal:
            qpos = (n - (qe - qs) + torch.arange(qe - qs, device=q.device))[:, None]
            logits = logits.masked_fill((torch.arange(n, device=q.device)[None, :] > qpos)[None], float("-inf"))
        out[qs:qe] = torch.einsum("hqn,nhd->qhd", torch.softmax(logits, -1), v)
        lse[qs:qe] = (torch.logsumexp(logits, -1) * 1.4426950408889634).transpose(0, 1)
    return out.to(q.dtype), lse


def _batch_attention_init(*, batch_size=6, num_qo_heads=8, num_kv_heads=2, head_dim=128, page_size=16, device="cuda", seed=0):
    import flashinfer_b200 as fi

    g = torch.Generator(device="cpu").manual_seed(seed)
    kv_lens = [int(x) for x in torch.randint(1, 5 * page_size, (batch_size,), generator=g)]
    q_lens = [1 if i % 2 == 0 else min(kv_lens[i], int(torch.randint(2, 40, (1,), generator=g))) for i in range(batch_size)]
    n_pages = [(n + page_size - 1) // page_size for n in kv_lens]
    total = sum(n_pages)
    kv_indptr = torch.tensor([0] + list(torch.tensor(n_pages).cumsum(0)), dtype=torch.int32)
    qo_indptr = torch.tensor([0] + list(torch.tensor(q_lens).cumsum(0)), dtype=torch.int32)
    kv_indices = torch.randperm(total + 2, generator=g)[:total].to(torch.int32)
    mk = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16).to(device)  # noqa: E731
    w = fi.BatchAttention("NHD", device=device)
    w.plan(qo_indptr, kv_indptr, kv_indices, torch.tensor(kv_lens, dtype=torch.int32), num_qo_heads, num_kv_heads, head_dim, head_dim, page_size,
           causal=True, q_data_type=torch.bfloat16, kv_data_type=torch.bfloat16)
    return {"self": w, "q": mk(int(qo_indptr[-1]), num_qo_heads, head_dim),
            "kv_cache": (mk(total + 2, page_size, num_kv_heads, head_dim), mk(total + 2, page_size, num_kv_heads, head_dim))}


_PAGED2 = ("num_pages", "page_size", "num_kv_heads", "head_dim")
batch_attention_trace = TraceTemplate(
    op_type="batch_attention", name_fmt="batch_attention_h{num_qo_heads}_kv{num_kv_heads}_d{head_dim}_ps{page_size}",
    axes=[Var("total_q"), Var("num_pages"), Var("len_indptr"), Var("num_kv_indices"), Var("batch_size"), Const("num_qo_heads", abbrev="h"),
          Const("num_kv_heads", abbrev="kv"), Const("head_dim", abbrev="d"), Const("page_size", abbrev="ps")],
    inputs=[Tensor("q", ("total_q", "num_qo_heads", "head_dim")), Tensor("k_cache", _PAGED2, param="kv_cache", tuple_idx=0),
            Tensor("v_cache", _PAGED2, param="kv_cache", tuple_idx=1), Tensor("qo_indptr", ("len_indptr",), "int32", param="self._qo_indptr_host"),
            Tensor("kv_indptr", ("len_indptr",), "int32", param="self._kv_indptr_host"),
            Tensor("kv_indices", ("num_kv_indices",), "int32", param="self._kv_indices_host"),
            Tensor("kv_len_arr", ("batch_size",), "int32", param="self._kv_len_host"), Scalar("causal", "bool", param="self._causal"),
            Scalar("sm_scale", param="self._sm_scale")],
    outputs=[Tensor("output", ("total_q", "num_qo_heads", "head_dim"), dtype_from="q"), Tensor("lse", ("total_q", "num_qo_heads"), dtype="float32")],
    reference=_batch_attention_reference, init=_batch_attention_init, tags=("attention", "prefill", "decode", "paged", "mixed-batch"),
    constraints=("len_indptr == batch_size + 1",), tolerance="bf16",
    description="Mixed prefill / decode batch over one paged KV cache (decode-sized requests on the decode kernel, the rest on the prefill kernel)",
    test_sizes={"num_qo_heads": 4, "num_kv_heads": 2, "head_dim": 32, "page_size": 8})

__all__ = [n for n in dir() if n.endswith("_trace")]


# ------------------------------------------------------------------ trtllm-style pre-routed bf16 MoE (packed routing words)
def _routed_moe_reference(topk_ids, hidden_states, gemm1_weights, gemm2_weights):
    """topk_ids [T, K] int32 = (expert_id << 16) | bf16 bits of the routing weight; gemm1 [E, 2I, H] = [up | gate] rows."""
    ids = (topk_ids >> 16).to(torch.int64)
    wts = (topk_ids & 0xFFFF).to(torch.int16).view(torch.bfloat16).to(torch.float32)
    t_, h = hidden_states.shape
    inter = gemm2_weights.shape[2]
    out = torch.zeros(t_, h, dtype=torch.float32, device=hidden_states.device)
    for e in range(gemm1_weights.shape[0]):
        tok, slot = torch.nonzero(ids == e, as_tuple=True)
        if tok.numel() == 0:
            continue
        hid = hidden_states[tok].to(torch.float32) @ gemm1_weights[e].to(torch.float32).t()
        act = torch.nn.functional.silu(hid[:, inter:]) * hid[:, :inter]
        out.index_add_(0, tok, (act @ gemm2_weights[e].to(torch.float32).t()) * wts[tok, slot][:, None])
    return out.to(hidden_states.dtype)


def _routed_moe_init(*, seq_len=64, num_experts=8, hidden_size=4096, intermediate_size=1024, top_k=2, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = (torch.randn(seq_len, hidden_size, generator=g) * 0.5).to(torch.bfloat16)
    w1 = (torch.randn(num_experts, 2 * intermediate_size, hidden_size, generator=g) / hidden_size ** 0.5).to(torch.bfloat16)
    w2 = (torch.randn(num_experts, hidden_size, intermediate_size, generator=g) / intermediate_size ** 0.5).to(torch.bfloat16)
    scales, ids = torch.topk(torch.softmax(torch.randn(seq_len, num_experts, generator=g), -1), top_k)
    packed = (ids.to(torch.int32) << 16) | (scales.to(torch.bfloat16).view(torch.int16).to(torch.int32) & 0xFFFF)
    return {"topk_ids": packed.to(device), "hidden_states": x.to(device), "gemm1_weights": w1.to(device), "gemm2_weights": w2.to(device),
            "num_experts": num_experts, "top_k": top_k, "n_group": None, "topk_group": None, "intermediate_size": intermediate_size,
            "local_expert_offset": 0, "local_num_experts": num_experts, "use_shuffled_weight": False, "weight_layout": 0}     # plain K-major weights


def _moe_first(got, expected, kwargs):
    out = got[0][0] if isinstance(got[0], (list, tuple)) else got[0]
    torch.testing.assert_close(out.float(), expected[0].float(), atol=3e-2, rtol=3e-2)


trtllm_bf16_routed_moe_trace = TraceTemplate(
    op_type="moe", name_fmt="trtllm_bf16_routed_moe_e{num_experts}_h{hidden_size}_i{intermediate_size}_k{top_k}",
    axes=[Var("seq_len"), Const("num_experts", abbrev="e"), Const("hidden_size", abbrev="h"), Const("intermediate_size", abbrev="i"),
          Const("top_k", abbrev="k"), Var("gemm1_rows")],
    inputs=[Tensor("topk_ids", ("seq_len", "top_k"), dtype="int32"), Tensor("hidden_states", ("seq_len", "hidden_size")),
            Tensor("gemm1_weights", ("num_experts", "gemm1_rows", "hidden_size")),
            Tensor("gemm2_weights", ("num_experts", "hidden_size", "intermediate_size"))],
    outputs=[Tensor("output", ("seq_len", "hidden_size"), dtype_from="hidden_states")], reference=_routed_moe_reference, init=_routed_moe_init,
    compare=_moe_first, tags=("moe", "bf16", "pre-routed"), constraints=("gemm1_rows == 2 * intermediate_size",),
    description="trtllm-style bf16 MoE with packed pre-computed routing ((expert << 16) | bf16 weight), SwiGLU experts",
    test_sizes={"num_experts": 4, "hidden_size": 64, "intermediate_size": 32, "top_k": 2})

__all__ = [n for n in dir() if n.endswith("_trace")]


# ------------------------------------------------------------------ XQA decode (page table + sequence lengths, reference xqa.py:155)
def _xqa_reference(q, k_cache, v_cache, page_table, seq_lens):
    """q [B, 1, Hq, D]; k_cache / v_cache [pages, page_size, Hkv, D] (NHD); page_table [B, max_pages]; seq_lens [B, 1]."""
    b, _, hq, d = q.shape
    page_size, hkv = k_cache.shape[1], k_cache.shape[2]
    g = hq // hkv
    out = torch.zeros(b, 1, hq, d, dtype=torch.float32, device=q.device)
    for i in range(b):
        n = int(seq_lens.reshape(-1)[i])
        pages = page_table[i, : (n + page_size - 1) // page_size].long()
        k = k_cache[pages].reshape(-1, hkv, d)[:n].to(torch.float32).repeat_interleave(g, dim=1)
        v = v_cache[pages].reshape(-1, hkv, d)[:n].to(torch.float32).repeat_interleave(g, dim=1)
        logits = torch.einsum("hd,nhd->hn", q[i, 0].to(torch.float32), k) / (d ** 0.5)
        out[i, 0] = torch.einsum("hn,nhd->hd", torch.softmax(logits, -1), v)
    return out.to(q.dtype)


def _xqa_init(*, batch_size=8, num_qo_heads=32, num_kv_heads=8, head_dim=128, page_size=16, max_pages=8, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    lens = torch.randint(1, max_pages * page_size + 1, (batch_size,), generator=g)
    total = batch_size * max_pages
    table = torch.randperm(total, generator=g).view(batch_size, max_pages).to(torch.int32)
    mk = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16).to(device)  # noqa: E731
    return {"q": mk(batch_size, 1, num_qo_heads, head_dim), "k_cache": mk(total, page_size, num_kv_heads, head_dim),
            "v_cache": mk(total, page_size, num_kv_heads, head_dim), "page_table": table.to(device),
            "seq_lens": lens.to(torch.int32).view(batch_size, 1).to(device),
            "output": torch.empty(batch_size, 1, num_qo_heads, head_dim, dtype=torch.bfloat16, device=device),
            "workspace_buffer": torch.zeros(16 << 20, dtype=torch.uint8, device=device)}


xqa_trace = TraceTemplate(
    op_type="xqa", name_fmt="xqa_h{num_qo_heads}_kv{num_kv_heads}_d{head_dim}_ps{page_size}",
    axes=[Var("batch_size"), Var("num_pages"), Var("max_pages"), Const("num_qo_heads", abbrev="h"), Const("num_kv_heads", abbrev="kv"),
          Const("head_dim", abbrev="d"), Const("page_size", abbrev="ps")],
    inputs=[Tensor("q", ("batch_size", "one", "num_qo_heads", "head_dim")), Tensor("k_cache", ("num_pages", "page_size", "num_kv_heads", "head_dim")),
            Tensor("v_cache", ("num_pages", "page_size", "num_kv_heads", "head_dim")), Tensor("page_table", ("batch_size", "max_pages"), dtype="int32"),
            Tensor("seq_lens", ("batch_size", "one"), dtype="int32")],
    outputs=[Tensor("out", ("batch_size", "one", "num_qo_heads", "head_dim"), dtype_from="q", param="output")],
    reference=_xqa_reference, init=_xqa_init, tags=("attention", "decode", "paged", "xqa"), constraints=("one == 1",), tolerance="bf16",
    description="XQA-style batch decode: per-request page table rows + sequence lengths (served by the tcgen05 paged decode kernel)",
    test_sizes={"num_qo_heads": 4, "num_kv_heads": 2, "head_dim": 32, "page_size": 8, "max_pages": 5})

__all__ = [n for n in dir() if n.endswith("_trace")]
