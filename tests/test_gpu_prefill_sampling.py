"""GPU tests: tcgen05 prefill kernel (ragged + paged) and the sampling kernels."""
import pytest
import torch

import flashinfer_b200 as fi
from flashinfer_b200 import reference, sampling
from helpers import make_paged

pytestmark = pytest.mark.gpu

PREFILL = [
    ([128], [128], 1, 1, False, torch.bfloat16, 0),
    ([256], [256], 2, 1, True, torch.bfloat16, 0),
    ([100, 300, 17], [100, 300, 17], 8, 2, True, torch.bfloat16, 0),
    ([33, 257], [500, 1000], 4, 4, True, torch.float16, 0),
    ([512], [2048], 8, 2, False, torch.bfloat16, 0),
    ([100, 300, 17], [150, 300, 400], 8, 2, True, torch.bfloat16, 16),
    ([64, 129], [1000, 129], 4, 1, True, torch.float16, 32),
    ([700], [700], 4, 2, True, torch.bfloat16, 8),
    ([1000], [1000], 32, 8, True, torch.bfloat16, 0),
]


@pytest.mark.parametrize("cfg", PREFILL, ids=lambda c: f"q{c[0][:2]}-kv{c[1][:2]}-h{c[2]}_{c[3]}-c{int(c[4])}-ps{c[6]}")
def test_batch_prefill(cfg):
    q_lens, kv_lens, hq, hkv, causal, dt, ps = cfg
    B = len(q_lens)
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    qo = torch.tensor([0] + torch.tensor(q_lens).cumsum(0).tolist(), dtype=torch.int32)
    q = torch.randn(sum(q_lens), hq, 128, device="cuda", dtype=dt)
    if ps == 0:
        kvi = torch.tensor([0] + torch.tensor(kv_lens).cumsum(0).tolist(), dtype=torch.int32)
        k = torch.randn(sum(kv_lens), hkv, 128, device="cuda", dtype=dt)
        v = torch.randn(sum(kv_lens), hkv, 128, device="cuda", dtype=dt)
        w = fi.BatchPrefillWithRaggedKVCacheWrapper(ws)
        w.plan(qo, kvi, hq, hkv, 128, causal=causal, q_data_type=dt)
        o, lse = w.run(q, k, v, return_lse=True)
        outs, lses = [], []
        for b in range(B):
            o_r, l_r = reference.attention_ref(q[qo[b]:qo[b + 1]], k[kvi[b]:kvi[b + 1]], v[kvi[b]:kvi[b + 1]], causal)
            outs.append(o_r)
            lses.append(l_r)
        o_ref, lse_ref = torch.cat(outs), torch.cat(lses)
    else:
        indptr, indices, last, kc, vc = make_paged(kv_lens, hkv, 128, ps, "NHD", dt, "cuda")
        w = fi.BatchPrefillWithPagedKVCacheWrapper(ws)
        w.plan(qo, indptr, indices, last, hq, hkv, 128, ps, causal=causal, q_data_type=dt)
        o, lse = w.run(q, (kc, vc), return_lse=True)
        o_ref, lse_ref = reference.batch_paged_attention_ref(q, qo, kc, vc, indptr, indices.cuda(), last, "NHD", causal)
    torch.testing.assert_close(o.float(), o_ref.float(), rtol=3e-2, atol=3e-2)
    torch.testing.assert_close(lse, lse_ref, rtol=2e-3, atol=2e-3)


def test_single_prefill_window_softcap():
    q = torch.randn(300, 4, 128, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(900, 2, 128, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(900, 2, 128, device="cuda", dtype=torch.bfloat16)
    o = fi.single_prefill_with_kv_cache(q, k, v, causal=True, window_left=200, logits_soft_cap=20.0)
    o_ref, _ = reference.attention_ref(q, k, v, True, None, 20.0, 200)
    torch.testing.assert_close(o.float(), o_ref.float(), rtol=3e-2, atol=3e-2)


def test_softmax_and_renorm():
    logits = torch.randn(7, 32000, device="cuda") * 3
    temp = torch.rand(7, device="cuda") + 0.5
    torch.testing.assert_close(sampling.softmax(logits, temp), torch.softmax(logits / temp[:, None], -1), rtol=1e-4, atol=1e-6)
    probs = torch.softmax(logits, -1)
    r = sampling.top_k_renorm_probs(probs, 50)
    assert ((r > 0).sum(-1) == 50).all()
    torch.testing.assert_close(r.sum(-1), torch.ones(7, device="cuda"), rtol=1e-4, atol=1e-4)
    r = sampling.top_p_renorm_probs(probs, 0.9)
    sp = probs.sort(-1, descending=True).values
    nkeep = ((sp.cumsum(-1) - sp) < 0.9).sum(-1)
    assert (((r > 0).sum(-1) - nkeep).abs() <= 1).all()
    m = sampling.top_k_mask_logits(logits, 10)
    assert (torch.isfinite(m).sum(-1) == 10).all()
    # rows that already carry -inf (banned tokens / grammar masks) must still be cut down to k entries
    pre = logits.clone()
    pre[:, ::3] = float("-inf")
    m = sampling.top_k_mask_logits(pre, 10)
    assert (torch.isfinite(m).sum(-1) == 10).all()
    kept = torch.where(torch.isfinite(m), m, torch.full_like(m, -1e30)).topk(10).values
    torch.testing.assert_close(kept, pre.topk(10).values)


def test_sampling_distributions():
    torch.manual_seed(0)
    V, n = 64, 20000
    p = torch.softmax(torch.randn(V, device="cuda") * 2, -1)
    probs = p[None].repeat(n, 1).contiguous()
    s = sampling.sampling_from_probs(probs, seed=1234, offset=0)
    freq = torch.bincount(s.long(), minlength=V).float() / n
    assert (freq - p).abs().max() < 0.02
    # top-k: only the k most likely tokens may appear, with renormalised frequencies
    s = sampling.top_k_sampling_from_probs(probs, 5, seed=7, offset=0)
    top = p.topk(5).indices
    assert set(s.unique().tolist()) <= set(top.tolist())
    pk = torch.zeros_like(p)
    pk[top] = p[top] / p[top].sum()
    freq = torch.bincount(s.long(), minlength=V).float() / n
    assert (freq - pk).abs().max() < 0.02
    # top-p
    s = sampling.top_p_sampling_from_probs(probs, 0.7, seed=9, offset=0)
    sp, si = p.sort(descending=True)
    keep = si[: int(((sp.cumsum(0) - sp) < 0.7).sum())]
    assert set(s.unique().tolist()) <= set(keep.tolist())
    # min-p
    s = sampling.min_p_sampling_from_probs(probs, 0.2, seed=11, offset=0)
    allowed = (p >= 0.2 * p.max()).nonzero().flatten()
    assert set(s.unique().tolist()) <= set(allowed.tolist())
    # determinism for a fixed (seed, offset)
    a = sampling.top_k_top_p_sampling_from_probs(probs, 10, 0.9, seed=5, offset=3)
    b = sampling.top_k_top_p_sampling_from_probs(probs, 10, 0.9, seed=5, offset=3)
    assert torch.equal(a, b)


def test_chain_speculative_sampling_accepts_identical_draft():
    B, n, V = 8, 4, 100
    target = torch.softmax(torch.randn(B, n + 1, V, device="cuda"), -1)
    draft = target[:, :n].clone()
    ids = torch.multinomial(draft.view(-1, V), 1).view(B, n).int()
    out, acc, emi = sampling.chain_speculative_sampling(draft, ids, target, seed=3, offset=0)
    assert torch.equal(out[:, :n], ids)  # q == p -> always accepted
    assert (out[:, n] >= 0).all() and (emi == n).all()


@pytest.mark.gpu
@pytest.mark.parametrize("B,V", [(1, 128256), (64, 128256), (7, 32000), (300, 1001), (2, 999999), (1024, 4096)])
def test_softmax_cluster_shapes(B, V):
    torch.manual_seed(B + V)
    logits = torch.randn(B, V, device="cuda") * 3
    temp = torch.rand(B, device="cuda") + 0.5
    got = sampling.softmax(logits, temp)
    ref = torch.softmax(logits / temp[:, None], -1)
    torch.testing.assert_close(got, ref, rtol=2e-4, atol=1e-7)
    got2 = sampling.softmax(logits, 0.7)
    torch.testing.assert_close(got2, torch.softmax(logits / 0.7, -1), rtol=2e-4, atol=1e-7)
