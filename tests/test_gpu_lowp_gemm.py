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
