"""GPU numerics tests: every CUDA kernel vs. the fp32 PyTorch oracle of the same op."""
import pytest
import torch

import flashinfer_b200 as fi
from flashinfer_b200 import activation, cascade, norm, page, reference, rope
from helpers import make_paged

pytestmark = pytest.mark.gpu
DT = [torch.float16, torch.bfloat16]


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("shape", [(1, 128), (7, 4096), (33, 8192), (3, 16384), (5, 24576)])
def test_rmsnorm(dtype, shape):
    x = torch.randn(shape, device="cuda", dtype=dtype)
    w = torch.randn(shape[-1], device="cuda", dtype=dtype)
    torch.testing.assert_close(norm.rmsnorm(x, w), reference.rmsnorm_ref(x, w), rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(norm.gemma_rmsnorm(x, w), reference.rmsnorm_ref(x, w, 1e-6, 1.0), rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("dtype", DT)
def test_rmsnorm_3d_strided(dtype):
    base = torch.randn(9, 12, 128, device="cuda", dtype=dtype)
    x = base[:, 2:10]  # non-contiguous heads slice
    w = torch.randn(128, device="cuda", dtype=dtype)
    torch.testing.assert_close(norm.rmsnorm(x, w), reference.rmsnorm_ref(x, w), rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("hidden", [512, 4096, 11008])
def test_fused_add_rmsnorm(dtype, hidden):
    x = torch.randn(19, hidden, device="cuda", dtype=dtype)
    r = torch.randn(19, hidden, device="cuda", dtype=dtype)
    w = torch.randn(hidden, device="cuda", dtype=dtype)
    y_ref, r_ref = reference.fused_add_rmsnorm_ref(x, r, w)
    norm.fused_add_rmsnorm(x, r, w)
    torch.testing.assert_close(r, r_ref, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(x, y_ref, rtol=2e-2, atol=2e-2)


def test_rmsnorm_quant_and_silu():
    x = torch.randn(11, 2048, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(2048, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(11, 2048, device="cuda", dtype=torch.float8_e4m3fn)
    norm.rmsnorm_quant(out, x, w, 0.05)
    ref = (reference.rmsnorm_ref(x.float(), w) / 0.05).clamp(-448, 448)
    assert (out.float() - ref).abs().max() <= 0.07 * ref.abs().max()
    y = norm.fused_rmsnorm_silu(x, w)
    torch.testing.assert_close(y.float(), torch.nn.functional.silu(reference.rmsnorm_ref(x.float(), w)), rtol=3e-2, atol=3e-2)
    g, b = torch.randn(2048, device="cuda"), torch.randn(2048, device="cuda")
    torch.testing.assert_close(norm.layernorm(x, g, b), reference.layernorm_ref(x, g, b), rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("d", [128, 14336])
def test_act_and_mul(dtype, d):
    x = torch.randn(13, 2 * d, device="cuda", dtype=dtype)
    torch.testing.assert_close(activation.silu_and_mul(x), reference.silu_and_mul_ref(x), rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(activation.gelu_and_mul(x), reference.gelu_and_mul_ref(x), rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(activation.gelu_tanh_and_mul(x), reference.gelu_and_mul_ref(x, "tanh"), rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("interleave", [False, True])
@pytest.mark.parametrize("rotary_dim", [128, 64])
def test_rope(interleave, rotary_dim):
    nnz, hq, hk, d = 50, 8, 2, 128
    q = torch.randn(nnz, hq, d, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(nnz, hk, d, device="cuda", dtype=torch.bfloat16)
    pos = torch.randint(0, 5000, (nnz,), device="cuda", dtype=torch.int32)
    qo, ko = rope.apply_rope_pos_ids(q, k, pos, rotary_dim=rotary_dim, interleave=interleave)
    torch.testing.assert_close(qo, reference.apply_rope_ref(q, pos, rotary_dim, interleave), rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(ko, reference.apply_rope_ref(k, pos, rotary_dim, interleave), rtol=2e-2, atol=2e-2)
    # llama3.1 in place + indptr/offsets form
    q2, k2 = q.clone(), k.clone()
    rope.apply_llama31_rope_pos_ids_inplace(q2, k2, pos, rotary_dim=rotary_dim, interleave=interleave)
    torch.testing.assert_close(q2, reference.apply_rope_ref(q, pos, rotary_dim, interleave, 8.0, 5e5, (1, 4, 8192)), rtol=2e-2, atol=2e-2)
    indptr = torch.tensor([0, 20, 50], device="cuda", dtype=torch.int32)
    offsets = torch.tensor([7, 1000], device="cuda", dtype=torch.int32)
    p2 = torch.cat([torch.arange(20) + 7, torch.arange(30) + 1000]).int().cuda()
    a = rope.apply_rope(q, k, indptr, offsets, rotary_dim=rotary_dim, interleave=interleave)[0]
    torch.testing.assert_close(a, reference.apply_rope_ref(q, p2, rotary_dim, interleave), rtol=2e-2, atol=2e-2)


def test_rope_cos_sin_cache_and_fp8():
    nnz, hq, hk, d = 33, 4, 1, 128
    q = torch.randn(nnz, hq * d, device="cuda", dtype=torch.float16)
    k = torch.randn(nnz, hk * d, device="cuda", dtype=torch.float16)
    pos = torch.randint(0, 300, (nnz,), device="cuda")
    inv = 1.0 / (1e4 ** (torch.arange(0, d, 2, device="cuda").float() / d))
    ang = torch.arange(300, device="cuda").float()[:, None] * inv[None]
    cache = torch.cat([ang.cos(), ang.sin()], -1)
    qo, ko = rope.apply_rope_with_cos_sin_cache(pos, q, k, d, cache, is_neox=True)
    torch.testing.assert_close(qo.view(nnz, hq, d), reference.apply_rope_ref(q.view(nnz, hq, d), pos), rtol=1e-2, atol=1e-2)
    q8, k8, _, _ = rope.rope_quantize_fp8(q.view(nnz, hq, d), k.view(nnz, hk, d), None, None, cache, pos)
    assert q8.dtype == torch.float8_e4m3fn
    torch.testing.assert_close(q8.float(), qo.view(nnz, hq, d).float(), rtol=0.08, atol=0.08)


@pytest.mark.parametrize("layout", ["NHD", "HND"])
def test_append_paged_kv_cache(layout):
    kv_lens = [20, 33, 1]
    indptr, indices, last, kc, vc = make_paged(kv_lens, 4, 128, 16, layout, torch.bfloat16, "cuda")
    append_indptr = torch.tensor([0, 5, 8, 9], dtype=torch.int32, device="cuda")
    seq = torch.tensor(kv_lens, dtype=torch.int32, device="cuda")
    bi, pos = page.get_batch_indices_positions(append_indptr, seq, 9)
    assert bi.tolist() == [0] * 5 + [1] * 3 + [2]
    assert pos.tolist() == [15, 16, 17, 18, 19, 30, 31, 32, 0]
    k = torch.randn(9, 4, 128, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(9, 4, 128, device="cuda", dtype=torch.bfloat16)
    page.append_paged_kv_cache(k, v, bi, pos, (kc, vc), indices.cuda(), indptr.cuda(), last.cuda(), layout)
    kk, vv = reference.gather_paged_kv(kc, vc, indices.cuda(), indptr, last, 1, layout)
    torch.testing.assert_close(kk[30:33], k[5:8])
    torch.testing.assert_close(vv[30:33], v[5:8])


@pytest.mark.parametrize("dtype", DT + [torch.float32])
def test_merge_states(dtype):
    n, h, d = 37, 8, 128
    va, vb = torch.randn(n, h, d, device="cuda", dtype=dtype), torch.randn(n, h, d, device="cuda", dtype=dtype)
    sa, sb = torch.randn(n, h, device="cuda") * 3, torch.randn(n, h, device="cuda") * 3
    v, s = cascade.merge_state(va, sa, vb, sb)
    vr, sr = reference.merge_state_ref(va, sa, vb, sb)
    torch.testing.assert_close(v, vr, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(s, sr, rtol=1e-3, atol=1e-3)
    vs, ss = torch.randn(n, 5, h, d, device="cuda", dtype=dtype), torch.randn(n, 5, h, device="cuda") * 3
    v2, s2 = cascade.merge_states(vs, ss)
    v2r, s2r = reference.merge_states_ref(vs, ss)
    torch.testing.assert_close(v2, v2r, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(s2, s2r, rtol=1e-3, atol=1e-3)
    mask = torch.rand(n, device="cuda") > 0.5
    va2, sa2 = va.clone(), sa.clone()
    cascade.merge_state_in_place(va2, sa2, vb, sb, mask)
    torch.testing.assert_close(va2[mask], vr[mask], rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(va2[~mask], va[~mask])


def test_fused_rmsnorm_silu_nvfp4_output_gpu():
    from flashinfer_b200.quantization.fp4 import e2m1_and_ufp8sf_scale_to_float

    torch.manual_seed(0)
    x, w = torch.randn(300, 512, device="cuda").bfloat16(), (1 + 0.1 * torch.randn(512, device="cuda")).bfloat16()
    q, sf = fi.norm.fused_rmsnorm_silu(x, w, 1e-6, out=torch.empty(300, 256, dtype=torch.uint8, device="cuda"))
    ref = fi.norm.fused_rmsnorm_silu(x, w, 1e-6).float()
    d = e2m1_and_ufp8sf_scale_to_float(q.view(torch.uint8), sf.view(torch.uint8), None, 16, 1, False)
    assert torch.nn.functional.cosine_similarity(d.flatten(), ref.flatten(), dim=0) > 0.99


def test_rope_quantize_fp8_nope_slices_and_mla_append():
    """No-rope slices are scaled / cast by the native kernel (rotary_dim = 0); the MLA flavour of the fused append writes
    (ckv, kpe) into the latent caches.  Oracle: the same ops on CPU tensors."""
    torch.manual_seed(0)
    T, H, dr, dn, page = 37, 16, 64, 512, 16
    q_rope, q_nope = torch.randn(T, H, dr).bfloat16(), torch.randn(T, H, dn).bfloat16()
    k_rope, k_nope = torch.randn(T, dr).bfloat16(), torch.randn(T, dn).bfloat16()
    pos = torch.arange(T, dtype=torch.int32)
    inv = 1.0 / (1e4 ** (torch.arange(0, dr, 2).float() / dr))
    ang = torch.arange(64).float()[:, None] * inv[None]
    cache = torch.cat([ang.cos(), ang.sin()], -1)
    cpu = rope.mla_rope_quantize_fp8(q_rope, k_rope, q_nope, k_nope, cache, pos, quant_scale_q=0.5, quant_scale_kv=2.0)
    gpu = rope.mla_rope_quantize_fp8(q_rope.cuda(), k_rope.cuda(), q_nope.cuda(), k_nope.cuda(), cache.cuda(), pos.cuda(),
                                     quant_scale_q=0.5, quant_scale_kv=2.0)
    for c, g in zip(cpu, gpu):
        assert g.dtype == torch.float8_e4m3fn and g.shape == c.shape
        torch.testing.assert_close(g.float().cpu(), c.float(), rtol=0.13, atol=0.05)  # at most one e4m3 step (rounding ties)
    # fused append, MLA layout: one request of T tokens
    n_pages = (T + page - 1) // page
    kv_indices = torch.randperm(n_pages).int()
    kv_indptr = torch.tensor([0, n_pages], dtype=torch.int32)
    bi = torch.zeros(T, dtype=torch.int32)
    outs = {}
    for dev in ("cpu", "cuda"):
        ckv = torch.zeros(n_pages, page, dn, dtype=torch.float8_e4m3fn, device=dev)
        kpe = torch.zeros(n_pages, page, dr, dtype=torch.float8_e4m3fn, device=dev)
        qr, qn = rope.rope_quantize_fp8_append_paged_kv_cache(
            q_rope.to(dev), k_rope.to(dev), q_nope.to(dev), k_nope.to(dev), None, cache.to(dev), pos.to(dev), (ckv, kpe),
            kv_indices.to(dev), kv_indptr.to(dev), bi.to(dev), pos.to(dev), quant_scale_kv=2.0, page_size=page)
        outs[dev] = (ckv.float().cpu(), kpe.float().cpu(), qr.float().cpu(), qn.float().cpu())
    for a, b in zip(outs["cpu"], outs["cuda"]):
        torch.testing.assert_close(b, a, rtol=0.13, atol=0.05)
    assert outs["cuda"][0].abs().sum() > 0 and outs["cuda"][1].abs().sum() > 0
