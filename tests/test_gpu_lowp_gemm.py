"""fp8 / mxfp8 / nvfp4 / mxfp4 tcgen05 GEMMs vs the fp32 de-quantised product (reference tests/gemm/test_mm_fp4.py,
test_bmm_fp8.py, test_mm_mxfp8.py use the same oracle with a cosine-similarity bound)."""
import os

import pytest
import torch

import flashinfer_b200 as fi
from flashinfer_b200.quantization.fp4 import e2m1_and_ufp8sf_scale_to_float

pytestmark = pytest.mark.gpu


def _to_fp8(x, dtype=torch.float8_e4m3fn):
    amax = x.abs().amax().clamp(min=1e-12)
    scale = torch.finfo(dtype).max / amax
    return (x * scale).clamp(-torch.finfo(dtype).max, torch.finfo(dtype).max).to(dtype), scale.float().reciprocal()


def _check(out, ref, tol=8e-3):
    err = (out.float() - ref).abs().max().item()
    assert err <= tol * max(ref.abs().max().item(), 1e-3), (err, ref.abs().max().item())


_BNS = [int(os.environ["FIB200_LOWP_BN_TEST"])] if "FIB200_LOWP_BN_TEST" in os.environ else [128, 256, 64, 192, 0]


@pytest.fixture(params=_BNS)
def bn(request):
    old = os.environ.get("FIB200_LOWP_BN")
    os.environ["FIB200_LOWP_BN"] = str(request.param)
    yield request.param
    if old is None:
        os.environ.pop("FIB200_LOWP_BN", None)
    else:
        os.environ["FIB200_LOWP_BN"] = old


@pytest.mark.parametrize("b,m,n,k", [(1, 48, 80, 64), (16, 48, 80, 64), (2, 300, 1000, 4096), (1, 1, 4096, 4096)])
@pytest.mark.parametrize("adt", [torch.float8_e4m3fn, torch.float8_e5m2])
def test_bmm_fp8(b, m, n, k, adt):
    A = torch.randn(b, m, k, device="cuda")
    W = torch.randn(b, n, k, device="cuda")
    a8, sa = _to_fp8(A, adt)
    w8, sw = _to_fp8(W)
    out = fi.bmm_fp8(a8, w8.transpose(-1, -2), sa, sw, torch.bfloat16)
    ref = torch.bmm(a8.float(), w8.float().transpose(-1, -2)) * sa * sw
    _check(out, ref)
    out2 = fi.mm_fp8(a8[0], w8[0].t(), sa * sw, torch.float16)
    _check(out2, ref[0])


@pytest.mark.parametrize("m,n,k", [(48, 256, 128), (128, 512, 4096), (300, 1184, 416), (1, 2048, 7168), (2000, 4096, 1024)])
def test_mm_mxfp8(m, n, k, bn):
    a = torch.randn(m, k, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(n, k, device="cuda", dtype=torch.bfloat16)
    aq, asf = fi.mxfp8_quantize(a)
    wq, wsf = fi.mxfp8_quantize(w)
    out = fi.mm_mxfp8(aq, wq.t(), asf, wsf, out_dtype=torch.bfloat16)
    ref = fi.mxfp8_dequantize_host(aq, asf) @ fi.mxfp8_dequantize_host(wq, wsf).t()
    _check(out, ref)
    # linear (2-D) scales take the on-the-fly swizzle path
    aq2, asf2 = fi.mxfp8_quantize(a, is_sf_swizzled_layout=False)
    out2 = fi.mm_mxfp8(aq2, wq.t(), asf2.view(m, k // 32), wsf, out_dtype=torch.bfloat16)
    _check(out2, ref)


@pytest.mark.parametrize("m,n,k", [(48, 256, 128), (128, 512, 4096), (300, 1184, 448), (1, 2048, 7168), (2000, 4096, 1024)])
@pytest.mark.parametrize("nv", [True, False])
def test_mm_fp4(m, n, k, nv, bn):
    a = torch.randn(m, k, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(n, k, device="cuda", dtype=torch.bfloat16)
    if nv:
        ga = (448 * 6) / a.float().abs().max()
        gw = (448 * 6) / w.float().abs().max()
        aq, asf = fi.nvfp4_quantize(a, ga)
        wq, wsf = fi.nvfp4_quantize(w, gw)
        alpha = 1.0 / (ga * gw)
        out = fi.mm_fp4(aq, wq.t(), asf, wsf, alpha, torch.bfloat16)
        ad = e2m1_and_ufp8sf_scale_to_float(aq, asf, ga, 16, 1, True)
        wd = e2m1_and_ufp8sf_scale_to_float(wq, wsf, gw, 16, 1, True)
    else:
        aq, asf = fi.mxfp4_quantize(a)
        wq, wsf = fi.mxfp4_quantize(w)
        out = fi.mm_fp4(aq, wq.t(), asf, wsf, None, torch.bfloat16, block_size=32, use_nvfp4=False)
        ad = fi.mxfp4_dequantize(aq, asf)
        wd = fi.mxfp4_dequantize(wq, wsf)
    _check(out, ad @ wd.t())


@pytest.mark.parametrize("m,n,k", [(16, 1024, 7168), (200, 512, 1024), (300, 1152, 2048), (1, 128, 128), (4096, 4096, 4096)])
@pytest.mark.parametrize("major", ["K", "MN"])
@pytest.mark.parametrize("gw_bn", [0, 32, 64, 128])
def test_gemm_fp8_groupwise(m, n, k, major, gw_bn):
    os.environ["FIB200_GW_BN"] = str(gw_bn)
    try:
        torch.manual_seed(0)
        a = torch.randn(m, k, device="cuda")
        w = torch.randn(n, k, device="cuda")
        sa = a.view(m, k // 128, 128).abs().amax(-1) / 448
        a8 = (a.view(m, k // 128, 128) / sa[..., None]).view(m, k).to(torch.float8_e4m3fn)
        sw = w.view(n // 128, 128, k // 128, 128).abs().amax((1, 3)) / 448
        w8 = (w.view(n // 128, 128, k // 128, 128) / sw[:, None, :, None]).view(n, k).to(torch.float8_e4m3fn)
        ref = (a8.float() * sa.repeat_interleave(128, 1)) @ (w8.float() * sw.repeat_interleave(128, 0).repeat_interleave(128, 1)).t()
        if major == "MN":
            sa_in, sw_in = sa.t().contiguous(), sw.t().contiguous()
        else:
            sa_in, sw_in = sa, sw
        out = fi.gemm_fp8_nt_groupwise(a8, w8, sa_in, sw_in, major, out_dtype=torch.bfloat16)
        _check(out, ref)
    finally:
        os.environ.pop("FIB200_GW_BN", None)


@pytest.mark.parametrize("m,n,k", [(48, 256, 4096), (512, 1024, 7168), (1, 512, 1024), (130, 448, 2048)])
@pytest.mark.parametrize("split", ["1", "2"])
def test_lowp_cluster_split_k(m, n, k, split):
    """Cluster split-K (two CTAs per tile, DSMEM reduction) for fp8 / mxfp8 / nvfp4."""
    os.environ["FIB200_LOWP_SPLIT"] = split
    try:
        torch.manual_seed(0)
        a = torch.randn(m, k, device="cuda", dtype=torch.bfloat16)
        w = torch.randn(n, k, device="cuda", dtype=torch.bfloat16)
        a8, sa = _to_fp8(a.float())
        w8, sw = _to_fp8(w.float())
        _check(fi.mm_fp8(a8, w8.t(), sa * sw, torch.bfloat16), (a8.float() @ w8.float().t()) * sa * sw)
        aq, asf = fi.mxfp8_quantize(a)
        wq, wsf = fi.mxfp8_quantize(w)
        _check(fi.mm_mxfp8(aq, wq.t(), asf, wsf), fi.mxfp8_dequantize_host(aq, asf) @ fi.mxfp8_dequantize_host(wq, wsf).t())
        g = torch.tensor(1.0, device="cuda")
        aq4, asf4 = fi.nvfp4_quantize(a, g)
        wq4, wsf4 = fi.nvfp4_quantize(w, g)
        ref = e2m1_and_ufp8sf_scale_to_float(aq4, asf4, g, 16, 1, True) @ e2m1_and_ufp8sf_scale_to_float(wq4, wsf4, g, 16, 1, True).t()
        _check(fi.mm_fp4(aq4, wq4.t(), asf4, wsf4, g, torch.bfloat16), ref)
    finally:
        os.environ.pop("FIB200_LOWP_SPLIT", None)


@pytest.mark.parametrize("m,n,k", [(512, 768, 1024), (300, 1184, 448), (2000, 4096, 1024), (4096, 4096, 4096)])
@pytest.mark.parametrize("bn2", ["0", "64", "128", "192"])
def test_lowp_cta_pair_kernel(m, n, k, bn2):
    """cta_group::2 kernel (256 x BN tile per CTA pair) for fp8 / mxfp8 / nvfp4 / mxfp4, forced on."""
    os.environ["FIB200_LOWP_2CTA"] = "1"
    os.environ["FIB200_LOWP_BN"] = bn2
    try:
        torch.manual_seed(0)
        a = torch.randn(m, k, device="cuda", dtype=torch.bfloat16)
        w = torch.randn(n, k, device="cuda", dtype=torch.bfloat16)
        a8, sa = _to_fp8(a.float())
        w8, sw = _to_fp8(w.float())
        _check(fi.mm_fp8(a8, w8.t(), sa * sw, torch.bfloat16), (a8.float() @ w8.float().t()) * sa * sw)
        A3 = torch.stack([a8, a8.flip(0)])
        W3 = torch.stack([w8, w8.flip(0)])
        o3 = fi.bmm_fp8(A3, W3.transpose(-1, -2), sa, sw, torch.bfloat16)
        _check(o3, torch.bmm(A3.float(), W3.float().transpose(-1, -2)) * sa * sw)
        aq, asf = fi.mxfp8_quantize(a)
        wq, wsf = fi.mxfp8_quantize(w)
        _check(fi.mm_mxfp8(aq, wq.t(), asf, wsf), fi.mxfp8_dequantize_host(aq, asf) @ fi.mxfp8_dequantize_host(wq, wsf).t())
        g = torch.tensor(1.0, device="cuda")
        aq4, asf4 = fi.nvfp4_quantize(a, g)
        wq4, wsf4 = fi.nvfp4_quantize(w, g)
        ref = e2m1_and_ufp8sf_scale_to_float(aq4, asf4, g, 16, 1, True) @ e2m1_and_ufp8sf_scale_to_float(wq4, wsf4, g, 16, 1, True).t()
        _check(fi.mm_fp4(aq4, wq4.t(), asf4, wsf4, g, torch.bfloat16), ref)
        aqm, asfm = fi.mxfp4_quantize(a)
        wqm, wsfm = fi.mxfp4_quantize(w)
        _check(fi.mm_fp4(aqm, wqm.t(), asfm, wsfm, None, torch.bfloat16, block_size=32, use_nvfp4=False),
               fi.mxfp4_dequantize(aqm, asfm) @ fi.mxfp4_dequantize(wqm, wsfm).t())
    finally:
        os.environ.pop("FIB200_LOWP_2CTA", None)
        os.environ.pop("FIB200_LOWP_BN", None)


def _q128(x):
    """1x128 group quantisation oracle -> (e4m3, fp32 scales, de-quantised fp32)."""
    g = x.float().view(x.shape[0], -1, 128)
    sc = g.abs().amax(-1).clamp_min(1e-10) / 448.0
    q = (g / sc[..., None]).view_as(x).to(torch.float8_e4m3fn)
    return q, sc, (q.float().view_as(g) * sc[..., None]).view_as(x)


def _qblk(w):
    E, N, K = w.shape
    blk = w.float().reshape(E, N // 128, 128, K // 128, 128)
    s = blk.abs().amax((2, 4)).clamp_min(1e-10) / 448.0
    q = (blk / s[:, :, None, :, None]).reshape(E, N, K).to(torch.float8_e4m3fn)
    dq = (q.float().reshape(E, N // 128, 128, K // 128, 128) * s[:, :, None, :, None]).reshape(E, N, K)
    return q, s, dq


def test_fp8_group_quantize_gpu():
    from flashinfer_b200.gemm.lowp import fp8_group_quantize

    torch.manual_seed(0)
    x = (torch.randn(300, 1024, device="cuda") * 3).bfloat16()
    q, sc = fp8_group_quantize(x)
    q_ref, sc_ref, _ = _q128(x)
    assert torch.allclose(sc, sc_ref, rtol=1e-5, atol=1e-9)
    assert (q.float() - q_ref.float()).abs().max().item() <= 32.0  # at most one e4m3 ulp at the top binade (rounding of 1/scale)
    assert ((q.float() - q_ref.float()) != 0).float().mean().item() < 0.02
    # gated + row_list + gather
    h = (torch.randn(64, 2 * 256, device="cuda")).bfloat16()
    rows = torch.tensor([5, -1, 0, 17, 33, 2], dtype=torch.int32, device="cuda")
    qg, sg = fp8_group_quantize(h, rows=64, gated=True, row_list=rows)
    act = h.float()[:, :256] * torch.nn.functional.silu(h.float()[:, 256:])
    for j, r in enumerate(rows.tolist()):
        if r < 0:
            continue
        _, s_ref, dq_ref = _q128(act[r:r + 1])
        assert torch.allclose(sg[r], s_ref[0], rtol=2e-2)
        got = qg[r].float().view(-1, 128) * sg[r][:, None]
        assert (got.flatten() - act[r]).abs().max().item() < 0.08 * act[r].abs().max().item() + 1e-3
    src = (torch.randn(10, 256, device="cuda")).bfloat16()
    lst = torch.tensor([3, 1, -1, 7, 2, 0], dtype=torch.int32, device="cuda")  # entry j reads token j // 2
    qq, ss = fp8_group_quantize(src, rows=8, row_list=lst, gather=True, list_div=2)
    for j, r in enumerate(lst.tolist()):
        if r >= 0:
            q_ref, s_ref, _ = _q128(src[j // 2:j // 2 + 1])
            assert torch.allclose(ss[r], s_ref[0], rtol=1e-5)


@pytest.mark.parametrize("E,N,K", [(4, 256, 512), (8, 1024, 1024), (3, 96, 256)])
def test_grouped_fp8_groupwise_gpu(E, N, K):
    """m-grouped contiguous fp8 GEMM with DeepSeek scales (native grouped mode) + the DeepGEMM-style wrappers."""
    from flashinfer_b200.gemm import batch_deepgemm_fp8_nt_groupwise, group_deepgemm_fp8_nt_groupwise

    torch.manual_seed(E + N)
    tiles = [2, 1, 3, 1, 2, 1, 1, 2][:E]
    m_idx = torch.cat([torch.full((t * 128,), e, dtype=torch.int32) for e, t in enumerate(tiles)]).cuda()
    M = m_idx.numel()
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(E, (N + 127) // 128 * 128, K, device="cuda") / K ** 0.5
    aq, a_s, a_dq = _q128(a)
    wq, w_s, w_dq = _qblk(w)
    wq, w_dq = wq[:, :N].contiguous(), w_dq[:, :N]
    ref = torch.stack([a_dq[i] @ w_dq[int(m_idx[i])].t() for i in range(0, M, 1)]) if M <= 512 else None
    if ref is None:
        ref = torch.empty(M, N, device="cuda")
        for e in range(E):
            sel = m_idx == e
            ref[sel] = a_dq[sel] @ w_dq[e].t()
    out = group_deepgemm_fp8_nt_groupwise(aq, wq, a_s, w_s, m_idx)
    err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-2, err
    if N % 128 == 0:
        # masked layout: group g has masked_m[g] live rows out of 256
        G = min(E, 4)
        masked = torch.tensor([256, 100, 0, 129][:G], dtype=torch.int32, device="cuda")
        a3 = torch.randn(G, 256, K, device="cuda")
        a3q, a3s, a3dq = _q128(a3.view(-1, K))
        o = batch_deepgemm_fp8_nt_groupwise(a3q.view(G, 256, K), wq[:G], a3s.view(G, 256, -1), w_s[:G], masked)
        for g in range(G):
            n = int(masked[g])
            if n:
                r = a3dq.view(G, 256, K)[g, :n] @ w_dq[g].t()
                assert (o[g, :n].float() - r).abs().max().item() / r.abs().max().item() < 1e-2


@pytest.mark.parametrize("block,swz", [(16, True), (16, False), (32, True)])
@pytest.mark.parametrize("with_res", [False, True])
def test_fused_rmsnorm_fp4quant_gpu(block, swz, with_res):
    """(add +) RMSNorm + FP4 quantisation in one kernel vs the two-kernel composition (norm kernel, then the quantiser)."""
    import flashinfer_b200 as fi
    from flashinfer_b200.quantization.fp4 import fp4_quantize

    torch.manual_seed(block + with_res)
    rows, h = 77, 4096
    x = torch.randn(rows, h, device="cuda").bfloat16()
    w = (1 + 0.1 * torch.randn(h, device="cuda")).bfloat16()
    gs = torch.tensor([2.0], device="cuda")
    if with_res:
        r0 = torch.randn(rows, h, device="cuda").bfloat16()
        r_fused = r0.clone()
        out = fi.add_rmsnorm_fp4quant(x, r_fused, w, global_scale=gs, eps=1e-5, block_size=block, is_sf_swizzled_layout=swz,
                                      output_both_sf_layouts=True)
        q, sf, sf_other = out
        xr, rr = x.clone(), r0.clone()
        fi.fused_add_rmsnorm(xr, rr, w, 1e-5)
        assert torch.equal(rr, r_fused)
        y = xr
    else:
        q, sf = fi.rmsnorm_fp4quant(x, w, global_scale=gs, eps=1e-5, block_size=block, is_sf_swizzled_layout=swz)
        y = fi.rmsnorm(x, w, 1e-5)
    q_ref, sf_ref = fp4_quantize(y, gs, sf_vec_size=block, sf_use_ue8m0=(block == 32), is_sf_swizzled_layout=swz)
    assert (q.view(torch.uint8) != q_ref.view(torch.uint8)).float().mean().item() < 5e-3
    assert (sf.view(torch.uint8).reshape(-1)[: sf_ref.numel()] != sf_ref.view(torch.uint8).reshape(-1)).float().mean().item() < 5e-3
    if with_res:
        _, sf_o_ref = fp4_quantize(y, gs, sf_vec_size=block, sf_use_ue8m0=(block == 32), is_sf_swizzled_layout=not swz)
        assert (sf_other.view(torch.uint8).reshape(-1)[: sf_o_ref.numel()] != sf_o_ref.view(torch.uint8).reshape(-1)).float().mean().item() < 5e-3
