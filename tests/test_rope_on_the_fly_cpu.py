"""pos_encoding_mode="ROPE_LLAMA": attention over UN-rotated keys with RoPE applied on the fly equals rotating q / k first and
attending in the plain mode (single, ragged, paged prefill, paged decode; NHD and HND pages)."""
import pytest
import torch

import flashinfer_b200 as fi


def _rot(x, pos, theta=1e4, scale=1.0):
    d = x.shape[-1]
    inv = theta ** (-torch.arange(0, d, 2).float() / d) / scale
    ang = pos.float()[:, None, None] * inv
    x1, x2 = x[..., : d // 2].float(), x[..., d // 2:].float()
    return torch.cat([x1 * ang.cos() - x2 * ang.sin(), x2 * ang.cos() + x1 * ang.sin()], -1).to(x.dtype)


def test_single_request_forms():
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(n, h, 64, generator=g).to(torch.bfloat16) for n, h in ((5, 4), (12, 2), (12, 2)))
    want = fi.single_prefill_with_kv_cache(_rot(q, torch.arange(7, 12)), _rot(k, torch.arange(12)), v, causal=True)
    got = fi.single_prefill_with_kv_cache(q, k, v, causal=True, pos_encoding_mode="ROPE_LLAMA")
    torch.testing.assert_close(got.float(), want.float(), atol=2e-2, rtol=2e-2)
    got = fi.single_prefill_with_kv_cache(q, k, v, causal=True, pos_encoding_mode="ROPE_LLAMA", rope_theta=5e5, rope_scale=2.0)
    want = fi.single_prefill_with_kv_cache(_rot(q, torch.arange(7, 12), 5e5, 2.0), _rot(k, torch.arange(12), 5e5, 2.0), v, causal=True)
    torch.testing.assert_close(got.float(), want.float(), atol=2e-2, rtol=2e-2)
    qd = q[0]
    want = fi.single_decode_with_kv_cache(_rot(qd[None], torch.tensor([11]))[0], _rot(k, torch.arange(12)), v)
    got = fi.single_decode_with_kv_cache(qd, k, v, pos_encoding_mode="ROPE_LLAMA")
    torch.testing.assert_close(got.float(), want.float(), atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("layout", ["NHD", "HND"])
def test_batched_wrappers(layout):
    g = torch.Generator().manual_seed(1)
    page_size, hq, hkv, d = 4, 4, 2, 64
    kv_lens, q_lens = [9, 4, 14], [3, 4, 1]
    per = [(n + page_size - 1) // page_size for n in kv_lens]
    ids = torch.randperm(sum(per) + 2, generator=g)[: sum(per)].int()
    indptr = torch.tensor([0] + list(torch.tensor(per).cumsum(0)), dtype=torch.int32)
    last = torch.tensor([(n - 1) % page_size + 1 for n in kv_lens], dtype=torch.int32)
    qo = torch.tensor([0] + list(torch.tensor(q_lens).cumsum(0)), dtype=torch.int32)
    shape = (sum(per) + 2, page_size, hkv, d) if layout == "NHD" else (sum(per) + 2, hkv, page_size, d)
    kc, vc = (torch.randn(*shape, generator=g).to(torch.bfloat16) for _ in range(2))
    q = torch.randn(sum(q_lens), hq, d, generator=g).to(torch.bfloat16)
    # oracle: rotate every request's keys by position, queries by (kv_len - q_len + i); run the plain mode on a rotated cache copy
    kc_rot = kc.clone()
    q_rot = q.clone()
    for r, n in enumerate(kv_lens):
        pages = ids[int(indptr[r]): int(indptr[r + 1])].long()
        rows = kc[pages] if layout == "NHD" else kc[pages].transpose(1, 2)
        flat = rows.reshape(-1, hkv, d)
        rot = _rot(flat, torch.arange(flat.shape[0])).reshape(rows.shape)
        kc_rot[pages] = rot if layout == "NHD" else rot.transpose(1, 2)
        qs, qe = int(qo[r]), int(qo[r + 1])
        q_rot[qs:qe] = _rot(q[qs:qe], torch.arange(n - (qe - qs), n))
    ws = torch.empty(1 << 20, dtype=torch.uint8)
    w = fi.BatchPrefillWithPagedKVCacheWrapper(ws, layout)
    w.plan(qo, indptr, ids, last, hq, hkv, d, page_size, causal=True, q_data_type=torch.bfloat16)
    want = w.run(q_rot, (kc_rot, vc))
    w.plan(qo, indptr, ids, last, hq, hkv, d, page_size, causal=True, q_data_type=torch.bfloat16, pos_encoding_mode="ROPE_LLAMA")
    got = w.run(q, (kc, vc))
    torch.testing.assert_close(got.float(), want.float(), atol=2e-2, rtol=2e-2)
    # decode (one query per request = its newest position)
    qd = torch.randn(3, hq, d, generator=g).to(torch.bfloat16)
    wd = fi.BatchDecodeWithPagedKVCacheWrapper(ws, layout)
    wd.plan(indptr, ids, last, hq, hkv, d, page_size, q_data_type=torch.bfloat16)
    want = wd.run(_rot(qd, torch.tensor(kv_lens) - 1), (kc_rot, vc))
    wd.plan(indptr, ids, last, hq, hkv, d, page_size, q_data_type=torch.bfloat16, pos_encoding_mode="ROPE_LLAMA")
    torch.testing.assert_close(wd.run(qd, (kc, vc)).float(), want.float(), atol=2e-2, rtol=2e-2)
    # ragged keys
    if layout == "NHD":
        kr = torch.randn(sum(kv_lens), hkv, d, generator=g).to(torch.bfloat16)
        vr = torch.randn(sum(kv_lens), hkv, d, generator=g).to(torch.bfloat16)
        kvp = torch.tensor([0] + list(torch.tensor(kv_lens).cumsum(0)), dtype=torch.int32)
        kr_rot = torch.cat([_rot(kr[int(kvp[r]): int(kvp[r + 1])], torch.arange(n)) for r, n in enumerate(kv_lens)])
        wr = fi.BatchPrefillWithRaggedKVCacheWrapper(ws, "NHD")
        wr.plan(qo, kvp, hq, hkv, d, causal=True, q_data_type=torch.bfloat16)
        want = wr.run(q_rot, kr_rot, vr)
        wr.plan(qo, kvp, hq, hkv, d, causal=True, q_data_type=torch.bfloat16, pos_encoding_mode="ROPE_LLAMA")
        torch.testing.assert_close(wr.run(q, kr, vr).float(), want.float(), atol=2e-2, rtol=2e-2)


def test_decode_alibi_is_served_by_the_prefill_kernel():
    from flashinfer_b200.utils import get_alibi_slopes

    g = torch.Generator().manual_seed(2)
    hq, hkv, d, n = 4, 2, 64, 11
    q, k, v = torch.randn(hq, d, generator=g).to(torch.bfloat16), torch.randn(n, hkv, d, generator=g).to(torch.bfloat16), torch.randn(n, hkv, d, generator=g).to(torch.bfloat16)
    slopes = get_alibi_slopes(hq).float()
    lg = torch.einsum("hd,nhd->hn", q.float(), k.float().repeat_interleave(2, 1)) / 8.0 + slopes[:, None] * (torch.arange(n) - (n - 1))[None, :]
    want = torch.einsum("hn,nhd->hd", torch.softmax(lg, -1), v.float().repeat_interleave(2, 1))
    got = fi.single_decode_with_kv_cache(q, k, v, pos_encoding_mode="ALIBI")
    torch.testing.assert_close(got.float(), want, atol=2e-2, rtol=2e-2)
    page_size = 4
    kc = torch.zeros(3, page_size, hkv, d, dtype=torch.bfloat16)
    vc = torch.zeros_like(kc)
    kc.view(-1, hkv, d)[:n], vc.view(-1, hkv, d)[:n] = k, v
    w = fi.BatchDecodeWithPagedKVCacheWrapper(torch.empty(1 << 20, dtype=torch.uint8), "NHD")
    w.plan(torch.tensor([0, 3], dtype=torch.int32), torch.arange(3, dtype=torch.int32), torch.tensor([3], dtype=torch.int32), hq, hkv, d, page_size,
           pos_encoding_mode="ALIBI", q_data_type=torch.bfloat16)
    torch.testing.assert_close(w.run(q[None], (kc, vc))[0].float(), want, atol=2e-2, rtol=2e-2)
