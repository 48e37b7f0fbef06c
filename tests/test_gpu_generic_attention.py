"""Configurations served by the catch-all attention kernel: head_dim 64 / 256, fp8 KV cache, custom masks
(reference tests/attention/test_batch_prefill_kernels.py, test_batch_decode_kernels.py, test_fp8_prefill.py grids)."""
import math

import pytest
import torch

import flashinfer_b200 as fi
from flashinfer_b200 import reference

pytestmark = pytest.mark.gpu


def _paged(kv_lens, page_size, hkv, d, dtype, layout="NHD"):
    n_pages = [(l + page_size - 1) // page_size for l in kv_lens]
    total = sum(n_pages)
    perm = torch.randperm(total)
    indptr = torch.tensor([0] + list(torch.tensor(n_pages).cumsum(0)), dtype=torch.int32)
    last = torch.tensor([(l - 1) % page_size + 1 for l in kv_lens], dtype=torch.int32)
    shape = (total, page_size, hkv, d) if layout == "NHD" else (total, hkv, page_size, d)
    k = (torch.randn(shape, device="cuda") * 0.5)
    v = (torch.randn(shape, device="cuda") * 0.5)
    return k.to(dtype), v.to(dtype), indptr, perm.int(), last


@pytest.mark.parametrize("d", [64, 256])
@pytest.mark.parametrize("layout", ["NHD", "HND"])
def test_decode_other_head_dims(d, layout):
    torch.manual_seed(0)
    B, hq, hkv, ps = 5, 8, 2, 16
    kv_lens = [1, 17, 300, 64, 1000]
    k, v, indptr, indices, last = _paged(kv_lens, ps, hkv, d, torch.bfloat16, layout)
    q = torch.randn(B, hq, d, device="cuda", dtype=torch.bfloat16)
    w = fi.BatchDecodeWithPagedKVCacheWrapper(torch.empty(32 << 20, dtype=torch.uint8, device="cuda"), layout)
    w.plan(indptr, indices, last, hq, hkv, d, ps, q_data_type=torch.bfloat16)
    out, lse = w.run(q, (k, v), return_lse=True)
    ref, lref = reference.batch_paged_attention_ref(q, torch.arange(B + 1, dtype=torch.int32), k, v, indptr, indices.cpu(), last,
                                                    layout, True, 1 / math.sqrt(d), 0.0, -1)
    assert (out.float() - ref.float()).abs().max() < 2e-2
    assert (lse - lref.to(lse.device)).abs().max() < 2e-2


@pytest.mark.parametrize("kv_dtype", [torch.float8_e4m3fn, torch.float8_e5m2])
def test_decode_fp8_kv(kv_dtype):
    torch.manual_seed(1)
    B, hq, hkv, d, ps = 4, 32, 8, 128, 16
    kv_lens = [33, 500, 128, 2049]
    k, v, indptr, indices, last = _paged(kv_lens, ps, hkv, d, kv_dtype)
    q = torch.randn(B, hq, d, device="cuda", dtype=torch.bfloat16)
    w = fi.BatchDecodeWithPagedKVCacheWrapper(torch.empty(32 << 20, dtype=torch.uint8, device="cuda"), "NHD")
    w.plan(indptr, indices, last, hq, hkv, d, ps, q_data_type=torch.bfloat16, kv_data_type=kv_dtype)
    out = w.run(q, (k, v))
    ref, _ = reference.batch_paged_attention_ref(q, torch.arange(B + 1, dtype=torch.int32), k.float().bfloat16(), v.float().bfloat16(),
                                                 indptr, indices.cpu(), last, "NHD", True, 1 / math.sqrt(d), 0.0, -1)
    # fp8 cache runs both MMAs as tcgen05 kind::f8f6f4 (Q and P are converted to e4m3): fp8-level error
    assert (out.float() - ref.float()).abs().max() < 4e-2
    cos = torch.nn.functional.cosine_similarity(out.float().flatten(), ref.float().flatten(), dim=0)
    assert cos > 0.995


@pytest.mark.parametrize("d", [64, 256])
@pytest.mark.parametrize("causal", [False, True])
def test_prefill_other_head_dims(d, causal):
    torch.manual_seed(2)
    hq, hkv = 8, 4
    qo = torch.tensor([0, 33, 33, 200], dtype=torch.int32)
    kv = torch.tensor([0, 64, 100, 400], dtype=torch.int32)
    q = torch.randn(200, hq, d, device="cuda", dtype=torch.float16)
    k = torch.randn(400, hkv, d, device="cuda", dtype=torch.float16)
    v = torch.randn(400, hkv, d, device="cuda", dtype=torch.float16)
    w = fi.BatchPrefillWithRaggedKVCacheWrapper(torch.empty(32 << 20, dtype=torch.uint8, device="cuda"))
    w.plan(qo, kv, hq, hkv, d, causal=causal, q_data_type=torch.float16)
    out, lse = w.run(q, k, v, return_lse=True)
    for b in range(3):
        qs, qe, ks, ke = int(qo[b]), int(qo[b + 1]), int(kv[b]), int(kv[b + 1])
        if qe == qs:
            continue
        ref, lref = reference.attention_ref(q[qs:qe], k[ks:ke], v[ks:ke], causal, 1 / math.sqrt(d))
        assert (out[qs:qe].float() - ref.float()).abs().max() < 2e-2
        assert (lse[qs:qe] - lref).abs().max() < 2e-2


def test_prefill_custom_mask_paged_and_single():
    torch.manual_seed(3)
    hq, hkv, d, ps = 4, 4, 128, 8
    qo_len, kv_len = 37, 90
    mask = torch.rand(qo_len, kv_len, device="cuda") > 0.4
    mask[:, 0] = True
    q = torch.randn(qo_len, hq, d, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(kv_len, hkv, d, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(kv_len, hkv, d, device="cuda", dtype=torch.bfloat16)
    ref, _ = reference.attention_ref(q, k, v, False, 1 / math.sqrt(d), custom_mask=mask)
    out = fi.single_prefill_with_kv_cache(q, k, v, custom_mask=mask)
    assert (out.float() - ref.float()).abs().max() < 2e-2
    # paged wrapper with the same mask
    n_pages = (kv_len + ps - 1) // ps
    kc = torch.zeros(n_pages, ps, hkv, d, device="cuda", dtype=torch.bfloat16)
    vc = torch.zeros_like(kc)
    kc.view(-1, hkv, d)[:kv_len] = k
    vc.view(-1, hkv, d)[:kv_len] = v
    w = fi.BatchPrefillWithPagedKVCacheWrapper(torch.empty(32 << 20, dtype=torch.uint8, device="cuda"))
    w.plan(torch.tensor([0, qo_len], dtype=torch.int32), torch.tensor([0, n_pages], dtype=torch.int32),
           torch.arange(n_pages, dtype=torch.int32), torch.tensor([(kv_len - 1) % ps + 1], dtype=torch.int32), hq, hkv, d, ps,
           custom_mask=mask.flatten(), q_data_type=torch.bfloat16)
    out2 = w.run(q, (kc, vc))
    assert (out2.float() - ref.float()).abs().max() < 2e-2


def test_single_decode_head_dim_64_softcap_window():
    torch.manual_seed(4)
    hq, hkv, d, L = 8, 8, 64, 777
    q = torch.randn(hq, d, device="cuda", dtype=torch.float16)
    k = torch.randn(L, hkv, d, device="cuda", dtype=torch.float16)
    v = torch.randn(L, hkv, d, device="cuda", dtype=torch.float16)
    out = fi.single_decode_with_kv_cache(q, k, v, window_left=200, logits_soft_cap=30.0)
    ref, _ = reference.attention_ref(q[None], k, v, True, 1 / math.sqrt(d), 30.0, 200)
    assert (out.float() - ref[0].float()).abs().max() < 2e-2


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_prefill_deepseek_head_dims_tcgen05(causal, dtype):
    """head_dim_qk 192 / head_dim_vo 128 (DeepSeek prefill) runs on the tcgen05 kernel."""
    torch.manual_seed(5)
    H = 16
    qo = torch.tensor([0, 100, 100, 700, 1500], dtype=torch.int32)
    q = torch.randn(1500, H, 192, device="cuda", dtype=dtype)
    k = torch.randn(1500, H, 192, device="cuda", dtype=dtype)
    v = torch.randn(1500, H, 128, device="cuda", dtype=dtype)
    w = fi.BatchPrefillWithRaggedKVCacheWrapper(torch.empty(64 << 20, dtype=torch.uint8, device="cuda"))
    w.plan(qo, qo, H, H, 192, head_dim_vo=128, causal=causal, q_data_type=dtype)
    out, lse = w.run(q, k, v, return_lse=True)
    assert out.shape == (1500, H, 128)
    for b in range(4):
        s, e = int(qo[b]), int(qo[b + 1])
        if s == e:
            continue
        ref, lref = reference.attention_ref(q[s:e], k[s:e], v[s:e], causal, 1 / math.sqrt(192))
        assert (out[s:e].float() - ref.float()).abs().max() < 2e-2
        assert (lse[s:e] - lref).abs().max() < 2e-2
    o1 = fi.single_prefill_with_kv_cache(q[:333], k[:333], v[:333], causal=causal)
    r1, _ = reference.attention_ref(q[:333], k[:333], v[:333], causal, 1 / math.sqrt(192))
    assert (o1.float() - r1.float()).abs().max() < 2e-2


@pytest.mark.parametrize("paged", [False, True])
def test_prefill_fp8_kv_stays_on_tensor_cores(paged):
    """fp8 KV + 16-bit Q prefill: KV is widened once and the tcgen05 kernel runs (BASELINE config: 8k prefill, fp8 KV)."""
    torch.manual_seed(6)
    hq, hkv, d, ps, L = 32, 8, 128, 16, 1024
    q = torch.randn(L, hq, d, device="cuda", dtype=torch.bfloat16)
    k = (torch.randn(L, hkv, d, device="cuda") * 0.5)
    v = (torch.randn(L, hkv, d, device="cuda") * 0.5)
    k8, v8 = k.to(torch.float8_e4m3fn), v.to(torch.float8_e4m3fn)
    ref, _ = reference.attention_ref(q, k8.float().bfloat16(), v8.float().bfloat16(), True, 1 / math.sqrt(d))
    ip = torch.tensor([0, L], dtype=torch.int32)
    if not paged:
        w = fi.BatchPrefillWithRaggedKVCacheWrapper(torch.empty(64 << 20, dtype=torch.uint8, device="cuda"))
        w.plan(ip, ip, hq, hkv, d, causal=True, q_data_type=torch.bfloat16, kv_data_type=torch.float8_e4m3fn)
        out = w.run(q, k8, v8)
    else:
        npg = L // ps
        perm = torch.randperm(npg + 5, device="cuda")[:npg]
        kc = torch.zeros(npg + 5, ps, hkv, d, device="cuda", dtype=torch.float8_e4m3fn)
        vc = torch.zeros_like(kc)
        kc[perm] = k8.view(npg, ps, hkv, d)
        vc[perm] = v8.view(npg, ps, hkv, d)
        w = fi.BatchPrefillWithPagedKVCacheWrapper(torch.empty(64 << 20, dtype=torch.uint8, device="cuda"))
        w.plan(ip, torch.tensor([0, npg], dtype=torch.int32), perm.int(), torch.tensor([ps], dtype=torch.int32), hq, hkv, d, ps,
               causal=True, q_data_type=torch.bfloat16, kv_data_type=torch.float8_e4m3fn)
        out = w.run(q, (kc, vc))
    assert (out.float() - ref.float()).abs().max() < 2e-2
