"""CPU tests of the C++ planners (no GPU needed): coverage, balance and merge bookkeeping."""
import pytest
import torch

import flashinfer_b200 as fi
from flashinfer_b200 import jit


def _plan(kv_lens, hkv, group, ps, num_ctas, q_lens=None, min_tiles=1):
    B = len(kv_lens)
    npages = [(l + ps - 1) // ps for l in kv_lens]
    indptr = torch.tensor([0] + torch.tensor(npages).cumsum(0).tolist(), dtype=torch.int32)
    kv = torch.tensor(kv_lens, dtype=torch.int32)
    qo = None
    if q_lens is not None:
        qo = torch.tensor([0] + torch.tensor(q_lens).cumsum(0).tolist(), dtype=torch.int32)
    max_segs = B * hkv + num_ctas + 1
    seg = torch.zeros(max_segs * 12, dtype=torch.int32)
    cta = torch.zeros(num_ctas + 1, dtype=torch.int32)
    mrg = torch.zeros((num_ctas + 1) * 8, dtype=torch.int32)
    counts = torch.zeros(8, dtype=torch.int64)
    jit.load("planner").call("decode_plan", indptr, kv, qo, B, hkv, group, ps, num_ctas, min_tiles, seg, max_segs, cta,
                             mrg, num_ctas + 1, counts)
    nseg, nmerge = int(counts[0]), int(counts[1])
    return seg.view(-1, 12)[:nseg], cta, mrg.view(-1, 8)[:nmerge], counts


def _tiles(l, ps):
    if l == 0:
        return 0
    if ps <= 128:
        tpt = (128 // ps) * ps
        return (l + tpt - 1) // tpt
    tpp = (ps + 127) // 128
    return (l // ps) * tpp + ((l % ps) + 127) // 128


@pytest.mark.parametrize("ps", [1, 16, 48, 128, 256, 200])
@pytest.mark.parametrize("kv_lens", [[1], [4096] * 64, [5, 0, 700, 128, 129, 33000], [17] * 300])
def test_decode_plan_covers_every_tile_once(ps, kv_lens):
    hkv, num_ctas = 8, 148
    seg, cta, mrg, counts = _plan(kv_lens, hkv, 4, ps, num_ctas)
    seen = {}
    for s in seg.tolist():
        b, h, t0, t1 = s[0], s[1], s[2], s[3]
        assert t1 > t0
        for t in range(t0, t1):
            assert (b, h, t) not in seen
            seen[(b, h, t)] = 1
        assert s[7] == kv_lens[b]
    expect = sum(_tiles(l, ps) for l in kv_lens) * hkv
    assert len(seen) == expect == int(counts[4])
    # CTA lists partition the segment list and are balanced to the quota
    quota = int(counts[5])
    assert cta[0] == 0 and cta[-1] == len(seg)
    for c in range(num_ctas):
        load = sum(s[3] - s[2] for s in seg[cta[c]:cta[c + 1]].tolist())
        # the LPT schedule may trade a few tiles for fewer segments / merges (cost model in planner.cpp)
        assert load <= max(2 * quota, quota + 8)
    # merge items reference consecutive slots of the same (req, head)
    slots = {}
    for s in seg.tolist():
        if s[4] >= 0:
            slots[s[4]] = (s[0], s[1])
    for m in mrg.tolist():
        for i in range(m[1]):
            assert slots[m[0] + i] == (m[5], m[4])
    assert len(slots) == int(counts[2])


def test_decode_plan_q_rows():
    seg, cta, mrg, counts = _plan([100, 200], 2, 4, 16, 8, q_lens=[3, 1])
    assert int(counts[3]) == 12
    assert seg[0, 5] == 0 and seg[0, 6] == 3
    assert seg[-1, 5] == 3 and seg[-1, 6] == 1


def test_prefill_plan_lpt_balance():
    B, H = 5, 8
    q_lens = [1000, 17, 4096, 300, 128]
    qo = torch.tensor([0] + torch.tensor(q_lens).cumsum(0).tolist(), dtype=torch.int32)
    kv = torch.tensor(q_lens, dtype=torch.int32)
    num_ctas = 148
    max_work = sum((q + 127) // 128 for q in q_lens) * H
    work = torch.zeros(max_work * 8, dtype=torch.int32)
    cta = torch.zeros(num_ctas + 1, dtype=torch.int32)
    counts = torch.zeros(4, dtype=torch.int64)
    jit.load("planner").call("prefill_plan", qo, kv, None, B, H, 128, 128, 1, -1, num_ctas, work, max_work, cta, counts)
    n = int(counts[0])
    assert n == max_work
    w = work.view(-1, 8)[:n]
    # every (req, q0, head) appears exactly once
    keys = {(a, b, d) for a, b, _, d, *_ in w.tolist()}
    assert len(keys) == n
    total_cost = 0
    for r in w.tolist():
        kv_hi = min(r[4], r[4] - r[5] + r[1] + r[2])
        total_cost += (kv_hi + 127) // 128 + 1
    assert int(counts[1]) <= total_cost / num_ctas * 1.5 + 34


def test_wrapper_plan_cpu_reference_roundtrip():
    from helpers import make_paged

    kv_lens = [37, 128, 300]
    indptr, indices, last, kc, vc = make_paged(kv_lens, 2, 128, 16)
    q = torch.randn(3, 8, 128)
    w = fi.BatchDecodeWithPagedKVCacheWrapper(torch.empty(16 << 20, dtype=torch.uint8))
    w.plan(indptr, indices, last, 8, 2, 128, 16, q_data_type=torch.float32)
    o, lse = w.run(q, (kc, vc), return_lse=True)
    for b in range(3):
        k, v = fi.reference.gather_paged_kv(kc, vc, indices, indptr, last, b)
        o1, l1 = fi.single_decode_with_kv_cache(q[b], k, v, return_lse=True)
        torch.testing.assert_close(o[b], o1)
        torch.testing.assert_close(lse[b], l1)


def _mla_plan_py(qo, kvp, kvl, causal, ctas, tile=32):
    """The planner as it was written in Python (mla/_core.py before the C++ port): the oracle of mla_plan."""
    rows = []
    for b in range(len(kvl)):
        ql = qo[b + 1] - qo[b]
        for i in range(ql):
            vis = kvl[b] - (ql - 1 - i) if causal else kvl[b]
            rows.append((qo[b] + i, kvp[b], max(vis, 0), kvp[b + 1] - kvp[b]))
    total = sum(r[2] for r in rows)
    chunk = max(4 * tile, -(-total // ctas))
    chunk = -(-chunk // tile) * tile
    target = max(ctas, len(rows))
    count = lambda c: sum(max(1, -(-r[2] // c)) for r in rows)  # noqa: E731
    if count(chunk) > target:
        lo, hi = chunk // tile, max(chunk // tile, -(-max((r[2] for r in rows), default=tile) // tile))
        while lo < hi:
            mid = (lo + hi) // 2
            if count(mid * tile) > target:
                lo = mid + 1
            else:
                hi = mid
        chunk = lo * tile
    kmax = max(1, max((-(-r[2] // chunk) for r in rows), default=1))
    work, parts = [], [1] * max(qo[-1], 1)
    for (qr, ps, vis, npg) in rows:
        nsp = max(1, -(-vis // chunk))
        parts[qr] = nsp
        for s in range(nsp):
            lo, hi = s * chunk, min(vis, (s + 1) * chunk)
            work.append([qr, ps, lo, max(hi, lo), vis, qr * kmax + s, max(npg, 1), (kmax << 16) | nsp])
    return work, parts, kmax, chunk


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("q_lens,kv_lens", [([1] * 64, [4096] * 64), ([1, 3, 1, 8], [5, 700, 0, 33000]), ([2] * 300, [17] * 300), ([1], [100000]),
                                            ([4, 4], [2, 1])])
def test_mla_plan_matches_python_oracle(causal, q_lens, kv_lens):
    page = 64
    qo = [0] + torch.tensor(q_lens).cumsum(0).tolist()
    npg = [(l + page - 1) // page for l in kv_lens]
    kvp = [0] + torch.tensor(npg).cumsum(0).tolist()
    for ctas in (74, 8):
        n_q = qo[-1]
        max_work = max(ctas, n_q, 1)
        work = torch.zeros(max_work * 8, dtype=torch.int32)
        parts = torch.zeros(max(n_q, 1), dtype=torch.int32)
        counts = torch.zeros(4, dtype=torch.int64)
        jit.load("planner").call("mla_plan", torch.tensor(qo), torch.tensor(kvp), torch.tensor(kv_lens), len(kv_lens), 1 if causal else 0, ctas, 32,
                                 work, max_work, parts, max(n_q, 1), counts)
        w_ref, p_ref, kmax, chunk = _mla_plan_py(qo, kvp, kv_lens, causal, ctas)
        assert int(counts[0]) == len(w_ref) and int(counts[1]) == kmax and int(counts[3]) == chunk
        assert work.view(-1, 8)[: len(w_ref)].tolist() == w_ref
        assert parts.tolist() == p_ref
        assert len(w_ref) <= max_work                                   # one wave of CTA pairs (or one item per query row)
