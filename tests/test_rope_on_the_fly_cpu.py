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


def test_pod_wrappers_pass_masks_and_positional_encodings_through():
    """POD with a custom prefill mask / ROPE_LLAMA equals the standalone prefill and decode calls with the same options."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from helpers import make_paged
    from flashinfer_b200.pod import BatchPODWithPagedKVCacheWrapper, PODWithPagedKVCacheWrapper

    torch.manual_seed(5)
    dt, ps, hq, hkv, d = torch.float32, 4, 4, 2, 64
    ws = torch.empty(1 << 22, dtype=torch.uint8)
    q_p, k_p, v_p = torch.randn(9, hq, d), torch.randn(13, hkv, d), torch.randn(13, hkv, d)
    indptr, indices, last, kc, vc = make_paged([7, 10], hkv, d, ps, "NHD", dt)
    q_d = torch.randn(2, hq, d)
    mask = torch.rand(9, 13) > 0.4
    mask[:, 0] = True
    # custom mask on the prefill side
    w = PODWithPagedKVCacheWrapper(ws)
    w.plan(indptr, indices, last, hq, hkv, d, ps, q_data_type=dt)
    o_p, o_d = w.run(q_p, k_p, v_p, q_d, (kc, vc), custom_mask_p=mask)
    torch.testing.assert_close(o_p, fi.single_prefill_with_kv_cache(q_p, k_p, v_p, custom_mask=mask), atol=1e-4, rtol=1e-4)
    dw = fi.BatchDecodeWithPagedKVCacheWrapper(ws, "NHD")
    dw.plan(indptr, indices, last, hq, hkv, d, ps, q_data_type=dt)
    torch.testing.assert_close(o_d, dw.run(q_d, (kc, vc)), atol=1e-4, rtol=1e-4)
    # ROPE_LLAMA on both sides (the decode side takes it from plan(), as in the reference)
    w.plan(indptr, indices, last, hq, hkv, d, ps, q_data_type=dt, pos_encoding_mode="ROPE_LLAMA")
    o_p, o_d = w.run(q_p, k_p, v_p, q_d, (kc, vc), causal_p=True, pos_encoding_mode_p="ROPE_LLAMA")
    torch.testing.assert_close(o_p, fi.single_prefill_with_kv_cache(q_p, k_p, v_p, causal=True, pos_encoding_mode="ROPE_LLAMA"), atol=1e-4, rtol=1e-4)
    dw.plan(indptr, indices, last, hq, hkv, d, ps, q_data_type=dt, pos_encoding_mode="ROPE_LLAMA")
    want_d = dw.run(q_d, (kc, vc))
    torch.testing.assert_close(o_d, want_d, atol=1e-4, rtol=1e-4)
    # batched POD: packed mask + rope from plan()
    qo_p = torch.tensor([0, 4, 9], dtype=torch.int32)
    ip, ii, il, kcp, vcp = make_paged([6, 9], hkv, d, ps, "NHD", dt)
    flat = torch.cat([(torch.rand(4, 6) > 0.3).flatten(), (torch.rand(5, 9) > 0.3).flatten()])
    flat[0], flat[24] = True, True
    wb = BatchPODWithPagedKVCacheWrapper(ws)
    wb.plan(qo_p, ip, ii, il, torch.arange(3, dtype=torch.int32), indptr, indices, last, hq, hkv, d, ps, q_data_type=dt, pos_encoding_mode="ROPE_LLAMA")
    o_p2, o_d2 = wb.run(q_p, (kcp, vcp), q_d, (kc, vc), custom_mask_p=flat)
    pw = fi.BatchPrefillWithPagedKVCacheWrapper(ws, "NHD")
    pw.plan(qo_p, ip, ii, il, hq, hkv, d, ps, custom_mask=flat, q_data_type=dt, pos_encoding_mode="ROPE_LLAMA")
    torch.testing.assert_close(o_p2, pw.run(q_p, (kcp, vcp)), atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(o_d2, want_d, atol=1e-4, rtol=1e-4)
    # causal after a mask re-plans back
    o_p3, _ = wb.run(q_p, (kcp, vcp), q_d, (kc, vc), causal_p=True)
    pw.plan(qo_p, ip, ii, il, hq, hkv, d, ps, causal=True, q_data_type=dt, pos_encoding_mode="ROPE_LLAMA")
    torch.testing.assert_close(o_p3, pw.run(q_p, (kcp, vcp)), atol=1e-4, rtol=1e-4)


def test_prefill_wrappers_honour_o_data_type():
    torch.manual_seed(1)
    hq, hkv, d = 4, 2, 64
    q, k, v = torch.randn(6, hq, d).bfloat16(), torch.randn(10, hkv, d).bfloat16(), torch.randn(10, hkv, d).bfloat16()
    ws = torch.empty(1 << 20, dtype=torch.uint8)
    w = fi.BatchPrefillWithRaggedKVCacheWrapper(ws, "NHD")
    qo, kv = torch.tensor([0, 6], dtype=torch.int32), torch.tensor([0, 10], dtype=torch.int32)
    w.plan(qo, kv, hq, hkv, d, causal=True, q_data_type=torch.bfloat16)
    base = w.run(q, k, v)
    w.plan(qo, kv, hq, hkv, d, causal=True, q_data_type=torch.bfloat16, o_data_type=torch.float32)
    o, lse = w.run(q, k, v, return_lse=True)
    assert o.dtype == torch.float32 and lse.shape == (6, hq)
    torch.testing.assert_close(o, base.float())
    buf = torch.empty(6, hq, d, dtype=torch.float16)
    w.plan(qo, kv, hq, hkv, d, causal=True, q_data_type=torch.bfloat16, o_data_type=torch.float16)
    assert w.run(q, k, v, out=buf) is buf
    torch.testing.assert_close(buf.float(), base.float(), atol=1e-2, rtol=1e-2)
    # planning again without o_data_type goes back to the query dtype
    w.plan(qo, kv, hq, hkv, d, causal=True, q_data_type=torch.bfloat16)
    assert w.run(q, k, v).dtype == torch.bfloat16
    kc, vc = torch.zeros(3, 4, hkv, d, dtype=torch.bfloat16), torch.zeros(3, 4, hkv, d, dtype=torch.bfloat16)
    kc.view(-1, hkv, d)[:10], vc.view(-1, hkv, d)[:10] = k, v
    pw = fi.BatchPrefillWithPagedKVCacheWrapper(ws, "NHD")
    pw.plan(qo, torch.tensor([0, 3], dtype=torch.int32), torch.arange(3, dtype=torch.int32), torch.tensor([2], dtype=torch.int32), hq, hkv, d, 4,
            causal=True, q_data_type=torch.bfloat16, o_data_type=torch.float32)
    o = pw.run(q, (kc, vc))
    assert o.dtype == torch.float32
    torch.testing.assert_close(o, base.float(), atol=2e-2, rtol=2e-2)


def test_deprecated_forward_replaces_the_planned_parameters():
    """begin_forward() + forward(causal=..., sm_scale=..., ...) of the reference era: the call-time parameters win, defaults included."""
    torch.manual_seed(3)
    hq, hkv, d, ps = 4, 2, 64, 4
    q, k, v = torch.randn(6, hq, d), torch.randn(10, hkv, d), torch.randn(10, hkv, d)
    ws = torch.empty(1 << 20, dtype=torch.uint8)
    qo, kv = torch.tensor([0, 6], dtype=torch.int32), torch.tensor([0, 10], dtype=torch.int32)
    w = fi.BatchPrefillWithRaggedKVCacheWrapper(ws, "NHD")
    w.begin_forward(qo, kv, hq, hkv, d, causal=True, q_data_type=torch.float32)
    torch.testing.assert_close(w.forward(q, k, v), fi.single_prefill_with_kv_cache(q, k, v, causal=False))       # forward's default: non-causal
    torch.testing.assert_close(w.forward(q, k, v, causal=True, sm_scale=0.05, logits_soft_cap=8.0),
                               fi.single_prefill_with_kv_cache(q, k, v, causal=True, sm_scale=0.05, logits_soft_cap=8.0))
    o, lse = w.forward_return_lse(q, k, v, causal=True, pos_encoding_mode="ROPE_LLAMA")
    want = fi.single_prefill_with_kv_cache(q, k, v, causal=True, pos_encoding_mode="ROPE_LLAMA", return_lse=True)
    torch.testing.assert_close(o, want[0])
    torch.testing.assert_close(lse, want[1])
    with pytest.raises(RuntimeError):
        fi.BatchPrefillWithRaggedKVCacheWrapper(ws, "NHD").forward(q, k, v)
    kc, vc = torch.zeros(3, ps, hkv, d), torch.zeros(3, ps, hkv, d)
    kc.view(-1, hkv, d)[:10], vc.view(-1, hkv, d)[:10] = k, v
    pw = fi.BatchPrefillWithPagedKVCacheWrapper(ws, "NHD")
    pw.begin_forward(qo, torch.tensor([0, 3], dtype=torch.int32), torch.arange(3, dtype=torch.int32), torch.tensor([2], dtype=torch.int32), hq, hkv, d, ps,
                     q_data_type=torch.float32)
    torch.testing.assert_close(pw.forward(q, (kc, vc), causal=True, window_left=3), fi.single_prefill_with_kv_cache(q, k, v, causal=True, window_left=3))
    dw = fi.BatchDecodeWithPagedKVCacheWrapper(ws, "NHD")
    dw.begin_forward(torch.tensor([0, 3], dtype=torch.int32), torch.arange(3, dtype=torch.int32), torch.tensor([2], dtype=torch.int32), hq, hkv, d, ps,
                     q_data_type=torch.float32, sm_scale=0.5)
    torch.testing.assert_close(dw.forward(q[:1], (kc, vc))[0], fi.single_decode_with_kv_cache(q[0], k, v))     # planned sm_scale replaced by the default
    o, lse = dw.forward_return_lse(q[:1], (kc, vc), sm_scale=0.05, window_left=4, v_scale=2.0)
    torch.testing.assert_close(o[0], 2.0 * fi.single_decode_with_kv_cache(q[0], k, v, sm_scale=0.05, window_left=4))
    assert lse.shape == (1, hq)
    from flashinfer_b200.sparse import BlockSparseAttentionWrapper

    bw = BlockSparseAttentionWrapper(ws)
    qs, ks, vs = torch.randn(8, hq, d), torch.randn(8, hkv, d), torch.randn(8, hkv, d)
    bw.begin_forward(torch.tensor([0, 2, 3], dtype=torch.int32), torch.tensor([0, 1, 1], dtype=torch.int32), 8, 8, 4, 4, hq, hkv, d, q_data_type=torch.float32,
                     sm_scale=0.3)
    dense = torch.tensor([[1, 1], [0, 1]], dtype=torch.bool).repeat_interleave(4, 0).repeat_interleave(4, 1)
    torch.testing.assert_close(bw.forward(qs, ks, vs, sm_scale=0.07), fi.single_prefill_with_kv_cache(qs, ks, vs, custom_mask=dense, sm_scale=0.07))


def test_reference_era_keyword_arguments():
    """Keyword arguments of the reference signatures that were missing: scales of single prefill / fmha_varlen, q_len_per_req of the
    decode wrapper, causal=False of the trtllm context entry point, the restating plan() arguments of the ragged wrapper."""
    torch.manual_seed(9)
    hq, hkv, d, ps = 4, 2, 64, 4
    q, k, v = torch.randn(6, hq, d), torch.randn(10, hkv, d), torch.randn(10, hkv, d)
    base = fi.single_prefill_with_kv_cache(q, k, v, causal=True, sm_scale=0.1)
    torch.testing.assert_close(fi.single_prefill_with_kv_cache(q, k, v, causal=True, sm_scale=0.05, k_scale=2.0, v_scale=3.0), 3.0 * base)
    with pytest.raises(NotImplementedError):
        fi.single_prefill_with_kv_cache(q, k, v, kv_cache_sf=(k, v))
    qo, kv = torch.tensor([0, 6], dtype=torch.int32), torch.tensor([0, 10], dtype=torch.int32)
    o = fi.prefill.fmha_varlen(q, k, v, qo, kv, causal=True, sm_scale=0.05, q_scale=4.0, k_scale=0.5, v_scale=3.0, o_scale=1.5)
    torch.testing.assert_close(o, 2.0 * base)
    ws = torch.empty(1 << 20, dtype=torch.uint8)
    w = fi.BatchPrefillWithRaggedKVCacheWrapper(ws, "NHD")
    w.plan(qo, kv, hq, hkv, d, causal=True, sm_scale=0.1, q_data_type=torch.float32, seq_lens=torch.tensor([10]), seq_lens_q=torch.tensor([6]),
           max_token_per_sequence=6, max_sequence_kv=10)
    torch.testing.assert_close(w.run(q, k, v, o_scale=0.5), 2.0 * base)
    with pytest.raises(NotImplementedError):
        w.plan(qo, kv, hq, hkv, d, v_indptr=kv)
    # decode wrapper: two query tokens per request == causal append of the newest two
    kc, vc = torch.zeros(3, ps, hkv, d), torch.zeros(3, ps, hkv, d)
    kc.view(-1, hkv, d)[:10], vc.view(-1, hkv, d)[:10] = k, v
    dw = fi.BatchDecodeWithPagedKVCacheWrapper(ws, "NHD")
    dw.plan(torch.tensor([0, 3], dtype=torch.int32), torch.arange(3, dtype=torch.int32), torch.tensor([2], dtype=torch.int32), hq, hkv, d, ps, q_data_type=torch.float32)
    got = dw.run(q[:2], (kc, vc), q_len_per_req=2, skip_softmax_threshold_scale_factor=1.0)
    torch.testing.assert_close(got, fi.single_prefill_with_kv_cache(q[:2], k, v, causal=True))
    torch.testing.assert_close(dw.run(q[:1], (kc, vc))[0], fi.single_decode_with_kv_cache(q[0], k, v))          # back to one token per request
    with pytest.raises(ValueError):
        dw.run(q[:3], (kc, vc), q_len_per_req=2)
    # trtllm context entry point, bidirectional
    tables, lens = torch.arange(3, dtype=torch.int32)[None], torch.tensor([10], dtype=torch.int32)
    o = fi.prefill.trtllm_batch_context_with_kv_cache(q, (kc, vc), ws, tables, lens, 6, 10, 0.1, 1.0, 1, qo, kv, kv_layout="NHD", causal=False)
    torch.testing.assert_close(o, fi.single_prefill_with_kv_cache(q, k, v, causal=False, sm_scale=0.1))
    with pytest.raises(ValueError):
        fi.prefill.trtllm_batch_context_with_kv_cache(q, (kc, vc), ws, tables, lens, 6, 10, 0.1, 1.0, 1, qo, kv, window_left=4, kv_layout="NHD", causal=False)
