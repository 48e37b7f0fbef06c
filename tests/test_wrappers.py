"""Wrapper-level tests (cascade, POD, block-sparse, BatchAttention, function-style APIs, sinks).  They run on CPU
through the fp32 oracle path and, on a B200, through the tcgen05 kernels (same python code above the kernel call)."""
import pytest
import torch

import flashinfer_b200 as fi
from flashinfer_b200 import reference
from flashinfer_b200.attention import BatchAttention
from flashinfer_b200.cascade import MultiLevelCascadeAttentionWrapper
from flashinfer_b200.decode import trtllm_batch_decode_with_kv_cache
from flashinfer_b200.pod import BatchPODWithPagedKVCacheWrapper, PODWithPagedKVCacheWrapper
from flashinfer_b200.sparse import BlockSparseAttentionWrapper
from helpers import make_paged

DEVICES = ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)]
D = 128


def _dt(device):
    return torch.float32 if device == "cpu" else torch.bfloat16


def _tol(device):
    return dict(rtol=1e-4, atol=1e-4) if device == "cpu" else dict(rtol=3e-2, atol=3e-2)


@pytest.mark.parametrize("device", DEVICES)
def test_multi_level_cascade_equals_full_attention(device):
    dt, ps, hq, hkv = _dt(device), 16, 8, 2
    # 2 requests share a 48-token prefix (3 pages), unique suffixes of 20 and 37 tokens
    prefix_pages, suf = 3, [20, 37]
    suf_pages = [(s + ps - 1) // ps for s in suf]
    total = prefix_pages + sum(suf_pages)
    kc = torch.randn(total, ps, hkv, D, device=device, dtype=dt)
    vc = torch.randn(total, ps, hkv, D, device=device, dtype=dt)
    q = torch.randn(2, hq, D, device=device, dtype=dt)
    ws = torch.empty(64 << 20, dtype=torch.uint8, device=device)
    w = MultiLevelCascadeAttentionWrapper(2, ws)
    qo0, qo1 = torch.tensor([0, 2], dtype=torch.int32), torch.tensor([0, 1, 2], dtype=torch.int32)
    kvp0 = torch.tensor([0, prefix_pages], dtype=torch.int32)
    kvp1 = torch.tensor([0, suf_pages[0], sum(suf_pages)], dtype=torch.int32)
    idx0 = torch.arange(prefix_pages, dtype=torch.int32)
    idx1 = torch.arange(prefix_pages, total, dtype=torch.int32)
    last0 = torch.tensor([ps], dtype=torch.int32)
    last1 = torch.tensor([(s - 1) % ps + 1 for s in suf], dtype=torch.int32)
    w.plan([qo0, qo1], [kvp0, kvp1], [idx0, idx1], [last0, last1], hq, hkv, D, ps, causal=True, q_data_type=dt)
    out = w.run(q, (kc, vc))
    off = prefix_pages
    for b in range(2):
        pages = list(range(prefix_pages)) + list(range(off, off + suf_pages[b]))
        off += suf_pages[b]
        k = kc[pages].reshape(-1, hkv, D)[: prefix_pages * ps + suf[b]]
        v = vc[pages].reshape(-1, hkv, D)[: prefix_pages * ps + suf[b]]
        # careful: suffix pages follow the prefix contiguously only after trimming the last suffix page
        k = torch.cat([kc[:prefix_pages].reshape(-1, hkv, D), kc[pages[prefix_pages:]].reshape(-1, hkv, D)[: suf[b]]])
        v = torch.cat([vc[:prefix_pages].reshape(-1, hkv, D), vc[pages[prefix_pages:]].reshape(-1, hkv, D)[: suf[b]]])
        o_ref, _ = reference.attention_ref(q[b : b + 1], k, v, True)
        torch.testing.assert_close(out[b : b + 1].float(), o_ref.float(), **_tol(device))


@pytest.mark.parametrize("device", DEVICES)
def test_pod_matches_separate_calls(device):
    dt, ps, hq, hkv = _dt(device), 16, 8, 2
    ws = torch.empty(128 << 20, dtype=torch.uint8, device=device)
    q_p = torch.randn(300, hq, D, device=device, dtype=dt)
    k_p = torch.randn(300, hkv, D, device=device, dtype=dt)
    v_p = torch.randn(300, hkv, D, device=device, dtype=dt)
    kv_lens = [100, 257, 31]
    indptr, indices, last, kc, vc = make_paged(kv_lens, hkv, D, ps, "NHD", dt, device)
    q_d = torch.randn(3, hq, D, device=device, dtype=dt)
    w = PODWithPagedKVCacheWrapper(ws)
    w.plan(indptr, indices, last, hq, hkv, D, ps, q_data_type=dt)
    o_p, o_d = w.run(q_p, k_p, v_p, q_d, (kc, vc), causal_p=True)
    ref_p, _ = reference.attention_ref(q_p, k_p, v_p, True)
    ref_d, _ = reference.batch_paged_attention_ref(q_d, torch.arange(4, dtype=torch.int32), kc, vc, indptr,
                                                   indices.to(device), last, "NHD", True)
    torch.testing.assert_close(o_p.float(), ref_p.float(), **_tol(device))
    torch.testing.assert_close(o_d.float(), ref_d.float(), **_tol(device))
    # batched POD
    qo_p = torch.tensor([0, 140, 300], dtype=torch.int32)
    ip, ii, il, kcp, vcp = make_paged([140, 200], hkv, D, ps, "NHD", dt, device)
    wb = BatchPODWithPagedKVCacheWrapper(ws)
    wb.plan(qo_p, ip, ii, il, torch.arange(4, dtype=torch.int32), indptr, indices, last, hq, hkv, D, ps, q_data_type=dt)
    (o_p2, o_d2) = wb.run(q_p, (kcp, vcp), q_d, (kc, vc), causal_p=True)
    ref_p2, _ = reference.batch_paged_attention_ref(q_p, qo_p, kcp, vcp, ip, ii.to(device), il, "NHD", True)
    torch.testing.assert_close(o_p2.float(), ref_p2.float(), **_tol(device))
    torch.testing.assert_close(o_d2.float(), ref_d.float(), **_tol(device))


@pytest.mark.parametrize("device", DEVICES)
def test_block_sparse_matches_dense_mask(device):
    dt, hq, hkv = _dt(device), 4, 2
    M, N, R, C = 512, 768, 128, 64
    mb, nb = M // R, N // C
    mask_blocks = torch.rand(mb, nb) > 0.5
    mask_blocks[:, 0] = True
    indptr = torch.tensor([0] + mask_blocks.sum(1).cumsum(0).tolist(), dtype=torch.int32)
    indices = torch.cat([torch.nonzero(mask_blocks[i]).flatten() for i in range(mb)]).int()
    q = torch.randn(M, hq, D, device=device, dtype=dt)
    k = torch.randn(N, hkv, D, device=device, dtype=dt)
    v = torch.randn(N, hkv, D, device=device, dtype=dt)
    w = BlockSparseAttentionWrapper(torch.empty(64 << 20, dtype=torch.uint8, device=device))
    w.plan(indptr, indices, M, N, R, C, hq, hkv, D, q_data_type=dt)
    o = w.run(q, k, v)
    dense = mask_blocks.repeat_interleave(R, 0).repeat_interleave(C, 1).to(device)
    o_ref, _ = reference.attention_ref(q, k, v, False, custom_mask=dense)
    torch.testing.assert_close(o.float(), o_ref.float(), **_tol(device))


@pytest.mark.parametrize("device", DEVICES)
def test_batch_attention_mixed_and_sinks(device):
    dt, ps, hq, hkv = _dt(device), 16, 8, 2
    kv_lens = [300, 64, 1000, 17]
    q_lens = [1, 64, 200, 1]
    indptr, indices, last, kc, vc = make_paged(kv_lens, hkv, D, ps, "NHD", dt, device)
    qo = torch.tensor([0] + torch.tensor(q_lens).cumsum(0).tolist(), dtype=torch.int32)
    q = torch.randn(sum(q_lens), hq, D, device=device, dtype=dt)
    ba = BatchAttention("NHD", device=device)
    ba.plan(qo, indptr, indices, torch.tensor(kv_lens, dtype=torch.int32), hq, hkv, D, D, ps, causal=True, q_data_type=dt,
            kv_data_type=dt)
    o, lse = ba.run(q, (kc, vc))
    o_ref, l_ref = reference.batch_paged_attention_ref(q, qo, kc, vc, indptr, indices.to(device), last, "NHD", True)
    torch.testing.assert_close(o.float(), o_ref.float(), **_tol(device))
    torch.testing.assert_close(lse, l_ref, rtol=2e-3, atol=2e-3)
    # attention sinks through the decode wrapper
    sinks = torch.randn(hq, device=device)
    w = fi.BatchDecodeWithPagedKVCacheWrapper(torch.empty(64 << 20, dtype=torch.uint8, device=device))
    w.plan(indptr, indices, last, hq, hkv, D, ps, q_data_type=dt)
    qd = torch.randn(4, hq, D, device=device, dtype=dt)
    o_s = w.run(qd, (kc, vc), sinks=sinks)
    for b in range(4):
        k, v = reference.gather_paged_kv(kc, vc, indices.to(device), indptr, last, b)
        o_r, _ = reference.attention_ref(qd[b : b + 1], k, v, False, sinks=sinks)
        torch.testing.assert_close(o_s[b : b + 1].float(), o_r.float(), **_tol(device))


@pytest.mark.parametrize("device", DEVICES)
def test_block_table_function_api(device):
    dt, ps, hq, hkv = _dt(device), 32, 8, 2
    B, max_pages = 3, 6
    seq = torch.tensor([100, 33, 192], dtype=torch.int32, device=device)
    kc = torch.randn(B * max_pages, hkv, ps, D, device=device, dtype=dt)  # HND
    vc = torch.randn(B * max_pages, hkv, ps, D, device=device, dtype=dt)
    bt = torch.randperm(B * max_pages).int().view(B, max_pages).to(device)
    q = torch.randn(B, hq, D, device=device, dtype=dt)
    ws = torch.empty(64 << 20, dtype=torch.uint8, device=device)
    o = trtllm_batch_decode_with_kv_cache(q, (kc, vc), ws, bt, seq, 192, bmm1_scale=D ** -0.5)
    for b in range(B):
        n = int(seq[b])
        k = kc[bt[b].long()].transpose(1, 2).reshape(-1, hkv, D)[:n]
        v = vc[bt[b].long()].transpose(1, 2).reshape(-1, hkv, D)[:n]
        o_r, _ = reference.attention_ref(q[b : b + 1], k, v, False)
        torch.testing.assert_close(o[b : b + 1].float(), o_r.float(), **_tol(device))


@pytest.mark.parametrize("device", DEVICES)
@pytest.mark.parametrize("layout", ["NHD", "HND"])
def test_rope_append_fused_matches_separate(device, layout):
    """apply_rope_append_paged_kv_cache == apply_rope_pos_ids + append_paged_kv_cache."""
    from flashinfer_b200 import page, rope

    torch.manual_seed(0)
    dt = torch.bfloat16 if device == "cuda" else torch.float32
    B, hq, hkv, d, ps = 5, 8, 2, 128, 16
    lens = torch.tensor([3, 17, 40, 16, 1])
    npg = (lens + ps - 1) // ps
    indptr = torch.zeros(B + 1, dtype=torch.int32)
    indptr[1:] = npg.cumsum(0)
    total = int(indptr[-1])
    indices = torch.randperm(total).int().to(device)
    shape = (total, ps, hkv, d) if layout == "NHD" else (total, hkv, ps, d)
    kc1, vc1 = torch.zeros(shape, dtype=dt, device=device), torch.zeros(shape, dtype=dt, device=device)
    kc2, vc2 = torch.zeros_like(kc1), torch.zeros_like(vc1)
    # append the LAST token of every request (decode step)
    pos = (lens - 1).int().to(device)
    bidx = torch.arange(B, dtype=torch.int32, device=device)
    q = torch.randn(B, hq, d, device=device, dtype=dt)
    k = torch.randn(B, hkv, d, device=device, dtype=dt)
    v = torch.randn(B, hkv, d, device=device, dtype=dt)
    q1, k1 = q.clone(), k.clone()
    rope.apply_llama31_rope_pos_ids_inplace(q1, k1, pos)
    page.append_paged_kv_cache(k1, v, bidx, pos, (kc1, vc1), indices, indptr.to(device), ((lens - 1) % ps + 1).int().to(device), layout)
    q2 = q.clone()
    rope.apply_rope_append_paged_kv_cache(q2, k.clone(), v, pos, bidx, (kc2, vc2), indices, indptr.to(device), layout,
                                          rope_scale=8, rope_theta=5e5, llama31=(1.0, 4.0, 8192.0))
    assert torch.equal(q1, q2)
    assert torch.equal(kc1, kc2)
    assert torch.equal(vc1, vc2)


@pytest.mark.gpu
def test_pod_runs_as_one_fused_kernel():
    """On the GPU the POD wrappers must take the single-kernel path (csrc/attention/pod_sm100.cu): the launches are
    counted inside pod_sm100.so, none in the stand-alone prefill / decode libraries, and the results match the
    two-stream composition (same kernel bodies, same plans)."""
    import ctypes

    from flashinfer_b200 import jit

    def count(name):
        f = jit.load(name)._dll.fib200_launch_count
        f.restype = ctypes.c_longlong
        return int(f())

    dt, ps, hq, hkv, device = torch.bfloat16, 16, 32, 8, "cuda"
    ws = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    torch.manual_seed(3)
    q_p = torch.randn(2048, hq, D, device=device, dtype=dt)
    k_p = torch.randn(2048, hkv, D, device=device, dtype=dt)
    v_p = torch.randn(2048, hkv, D, device=device, dtype=dt)
    kv_lens = [1000 + 37 * i for i in range(24)]
    indptr, indices, last, kc, vc = make_paged(kv_lens, hkv, D, ps, "NHD", dt, device)
    q_d = torch.randn(len(kv_lens), hq, D, device=device, dtype=dt)
    w = PODWithPagedKVCacheWrapper(ws)
    w.plan(indptr, indices, last, hq, hkv, D, ps, q_data_type=dt)
    w.run(q_p, k_p, v_p, q_d, (kc, vc), causal_p=True)  # warm-up (loads the libraries)
    c_pod, c_p, c_d = count("pod_sm100"), count("prefill_sm100"), count("decode_sm100")
    o_p, o_d = w.run(q_p, k_p, v_p, q_d, (kc, vc), causal_p=True)
    torch.cuda.synchronize()
    assert 1 <= count("pod_sm100") - c_pod <= 2  # fused kernel (+ the decode split-KV merge when the planner splits)
    assert count("prefill_sm100") == c_p and count("decode_sm100") == c_d
    w._fused = False
    r_p, r_d = w.run(q_p, k_p, v_p, q_d, (kc, vc), causal_p=True)
    torch.cuda.synchronize()
    torch.testing.assert_close(o_p.float(), r_p.float(), rtol=0, atol=1e-2)
    torch.testing.assert_close(o_d.float(), r_d.float(), rtol=0, atol=1e-2)
    ref_p, _ = reference.attention_ref(q_p[:256], k_p[:256], v_p[:256], True)
    torch.testing.assert_close(o_p[:256].float(), ref_p.float(), rtol=3e-2, atol=3e-2)


@pytest.mark.parametrize("device", DEVICES)
def test_decode_nvfp4_kv_cache(device):
    """NVFP4 KV cache (packed e2m1 + per-16 UE4M3 scales in the cache layout, global scales through k_scale / v_scale)."""
    from flashinfer_b200.decode import BatchDecodeWithPagedKVCacheWrapper
    from flashinfer_b200.quantization.fp4 import nvfp4_dequantize_paged_kv_cache, nvfp4_quantize_paged_kv_cache

    dt, ps, hq, hkv = _dt(device), 16, 8, 2
    if device == "cpu":
        dt = torch.bfloat16
    kv_lens = [100, 33, 257]
    indptr, indices, last, kc, vc = make_paged(kv_lens, hkv, D, ps, "NHD", dt, device)
    q = torch.randn(len(kv_lens), hq, D, device=device, dtype=dt)
    (kq, vq), (ksf, vsf), kg, vg = nvfp4_quantize_paged_kv_cache(kc, vc, "NHD")
    assert kq.dtype == torch.uint8 and kq.shape[-1] == D // 2 and ksf.shape[-1] == D // 16
    w = BatchDecodeWithPagedKVCacheWrapper(torch.empty(32 << 20, dtype=torch.uint8, device=device), "NHD")
    w.plan(indptr, indices, last, hq, hkv, D, ps, q_data_type=dt)
    o = w.run(q, (kq, vq), kv_cache_sf=(ksf, vsf), k_scale=kg, v_scale=vg)
    kd = (nvfp4_dequantize_paged_kv_cache(kq, ksf, torch.float32) * kg).to(dt)
    vd = (nvfp4_dequantize_paged_kv_cache(vq, vsf, torch.float32) * vg).to(dt)
    assert (kd.float() - kc.float()).abs().mean().item() < 0.12 * kc.float().abs().mean().item()  # fp4 grid
    ref, _ = reference.batch_paged_attention_ref(q, torch.arange(len(kv_lens) + 1, dtype=torch.int32), kd, vd, indptr,
                                                 indices.to(device), last, "NHD", True)
    torch.testing.assert_close(o.float(), ref.float(), rtol=5e-2, atol=5e-2)
    with pytest.raises(ValueError):
        w.run(q, (kq, vq))
