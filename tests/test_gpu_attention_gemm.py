"""GPU tests of the tcgen05 kernels (GEMM, paged decode) against fp32 PyTorch references."""
import pytest
import torch

import flashinfer_b200 as fi
from flashinfer_b200 import reference
from flashinfer_b200.gemm import linear
from helpers import make_paged

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mnk", [(128, 128, 64), (300, 1000, 520), (1, 4096, 4096), (16, 6144, 4096), (64, 4096, 14336),
                                 (100, 384, 2048), (1024, 4096, 4096), (129, 136, 72),
                                 # decode shapes of the cluster split-K path with the A operand multicast across N-tile pairs
                                 (64, 4096, 4096), (64, 6144, 4096), (33, 4096, 14336), (64, 28672, 4096), (7, 1024, 2048)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_linear(mnk, dtype):
    m, n, k = mnk
    x = torch.randn(m, k, device="cuda", dtype=dtype)
    w = torch.randn(n, k, device="cuda", dtype=dtype) / k ** 0.5
    b = torch.randn(n, device="cuda", dtype=dtype)
    ref = x.float() @ w.float().t()
    torch.testing.assert_close(linear(x, w).float(), ref, rtol=2e-2, atol=3e-2)
    torch.testing.assert_close(linear(x, w, b).float(), ref + b.float(), rtol=2e-2, atol=3e-2)
    # reference-style API: b is [k, n] column-major
    torch.testing.assert_close(fi.mm_bf16(x, w.t(), out_dtype=dtype).float(), ref, rtol=2e-2, atol=3e-2)


DECODE_CFGS = [
    ([128], 4, 1, 16, "NHD", torch.bfloat16),
    ([37, 128, 300], 8, 2, 16, "NHD", torch.bfloat16),
    ([1, 17, 1000, 4096], 32, 8, 16, "NHD", torch.bfloat16),
    ([513, 64], 32, 8, 16, "HND", torch.float16),
    ([700, 90, 5], 16, 16, 32, "NHD", torch.bfloat16),
    ([700, 90, 5], 32, 4, 8, "HND", torch.bfloat16),
    ([333], 8, 8, 128, "NHD", torch.float16),
    ([1000, 33], 8, 1, 256, "NHD", torch.bfloat16),
    ([77, 200], 8, 2, 1, "NHD", torch.bfloat16),
    ([77, 200], 8, 2, 2, "HND", torch.bfloat16),
    ([100, 250], 8, 2, 48, "NHD", torch.bfloat16),
    ([2048] * 40, 32, 8, 16, "NHD", torch.bfloat16),
]


@pytest.mark.parametrize("cfg", DECODE_CFGS, ids=lambda c: f"kv{c[0][:2]}x{len(c[0])}-h{c[1]}_{c[2]}-ps{c[3]}-{c[4]}")
def test_batch_decode_paged(cfg):
    kv_lens, hq, hkv, ps, layout, dt = cfg
    B = len(kv_lens)
    indptr, indices, last, kc, vc = make_paged(kv_lens, hkv, 128, ps, layout, dt, "cuda")
    q = torch.randn(B, hq, 128, device="cuda", dtype=dt)
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    w = fi.BatchDecodeWithPagedKVCacheWrapper(ws, layout)
    w.plan(indptr, indices, last, hq, hkv, 128, ps, q_data_type=dt)
    o, lse = w.run(q, (kc, vc), return_lse=True)
    qo = torch.arange(B + 1, dtype=torch.int32)
    o_ref, lse_ref = reference.batch_paged_attention_ref(q, qo, kc, vc, indptr, indices.cuda(), last, layout, True)
    torch.testing.assert_close(o.float(), o_ref.float(), rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(lse, lse_ref, rtol=1e-3, atol=1e-3)


def test_decode_window_softcap_and_spec_qlen():
    kv_lens, hq, hkv, ps = [300, 1000, 64], 8, 2, 16
    indptr, indices, last, kc, vc = make_paged(kv_lens, hkv, 128, ps, "NHD", torch.bfloat16, "cuda")
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    q = torch.randn(3, hq, 128, device="cuda", dtype=torch.bfloat16)
    w = fi.BatchDecodeWithPagedKVCacheWrapper(ws)
    w.plan(indptr, indices, last, hq, hkv, 128, ps, q_data_type=torch.bfloat16, window_left=100, logits_soft_cap=30.0)
    o = w.run(q, (kc, vc))
    qo = torch.arange(4, dtype=torch.int32)
    o_ref, _ = reference.batch_paged_attention_ref(q, qo, kc, vc, indptr, indices.cuda(), last, "NHD", True, None, 30.0, 100)
    torch.testing.assert_close(o.float(), o_ref.float(), rtol=2e-2, atol=2e-2)
    # speculative decode: q_len = 3, 1, 4 tokens per request, causal inside the new tokens
    qo = torch.tensor([0, 3, 4, 8], dtype=torch.int32)
    q = torch.randn(8, hq, 128, device="cuda", dtype=torch.bfloat16)
    w.plan(indptr, indices, last, hq, hkv, 128, ps, q_data_type=torch.bfloat16, qo_indptr=qo)
    o, lse = w.run(q, (kc, vc), return_lse=True)
    o_ref, lse_ref = reference.batch_paged_attention_ref(q, qo, kc, vc, indptr, indices.cuda(), last, "NHD", True)
    torch.testing.assert_close(o.float(), o_ref.float(), rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(lse, lse_ref, rtol=1e-3, atol=1e-3)


def test_single_decode_and_cuda_graph():
    q = torch.randn(32, 128, device="cuda", dtype=torch.float16)
    k = torch.randn(777, 8, 128, device="cuda", dtype=torch.float16)
    v = torch.randn(777, 8, 128, device="cuda", dtype=torch.float16)
    o = fi.single_decode_with_kv_cache(q, k, v)
    o_ref, _ = reference.attention_ref(q[None], k, v)
    torch.testing.assert_close(o.float(), o_ref[0].float(), rtol=2e-2, atol=2e-2)
    # CUDA-graph capture of run()
    kv_lens = [500] * 8
    indptr, indices, last, kc, vc = make_paged(kv_lens, 8, 128, 16, "NHD", torch.bfloat16, "cuda")
    qq = torch.randn(8, 32, 128, device="cuda", dtype=torch.bfloat16)
    out = torch.empty_like(qq)
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    w = fi.BatchDecodeWithPagedKVCacheWrapper(ws)
    w.plan(indptr, indices, last, 32, 8, 128, 16, q_data_type=torch.bfloat16)
    w.run(qq, (kc, vc), out=out)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        w.run(qq, (kc, vc), out=out)
    ref = out.clone()
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    torch.testing.assert_close(out, ref)


@pytest.mark.parametrize("mnk", [(64, 28672, 4096), (1, 1024, 512), (200, 2048, 1024), (1024, 4096, 2048), (16, 96, 256)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_linear_gated_silu(mnk, dtype):
    """GEMM with the SwiGLU epilogue (row-interleaved gate / up weights) vs silu_and_mul(linear(x, [Wg; Wu]))."""
    from flashinfer_b200.gemm import interleave_gate_up, linear_gated_silu

    m, n2, k = mnk
    torch.manual_seed(m + k)
    x = torch.randn(m, k, device="cuda", dtype=dtype)
    w = torch.randn(n2, k, device="cuda", dtype=dtype) / k ** 0.5
    h = x.float() @ w.float().t()
    ref = torch.nn.functional.silu(h[:, : n2 // 2]) * h[:, n2 // 2:]
    out = linear_gated_silu(x, interleave_gate_up(w))
    assert out.shape == (m, n2 // 2)
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("m,n,k,sms", [(4096, 4096, 1024, 32), (300, 1000, 512, 7), (64, 4096, 4096, 40), (2048, 8192, 2048, None)])
def test_sm_constrained_gemm(m, n, k, sms):
    """reference flashinfer/triton/sm_constraint_gemm.py gemm_persistent(num_sms=...): same result on a capped persistent grid."""
    from flashinfer_b200.triton import gemm_persistent

    torch.manual_seed(0)
    a = (torch.randn(m, k, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(n, k, device="cuda") / k ** 0.5).bfloat16()
    ref = a.float() @ w.float().t()
    out = gemm_persistent(a, w.t(), num_sms=sms)
    assert (out.float() - ref).abs().max() / ref.abs().max() < 1e-2
    c = torch.randn(m, n, device="cuda").bfloat16()
    c0 = c.clone()
    gemm_persistent(a, w.t(), c, alpha=0.5, beta=2.0, num_sms=sms)
    ref2 = 0.5 * ref + 2.0 * c0.float()
    assert (c.float() - ref2).abs().max() / ref2.abs().max() < 1e-2
