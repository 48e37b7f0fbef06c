"""flashinfer.grouped_mm entry points in the reference's conventions (grouped_mm/core.py): per-tensor FP8 with alpha, block-scaled FP4
with 128x4-swizzled scales, against a per-expert loop on the de-quantised operands."""
import pytest
import torch

import flashinfer_b200 as fi
from flashinfer_b200.quantization.fp4 import fp4_quantize


def _loop(ad, bd, indptr):
    out = torch.zeros(ad.shape[0], bd.shape[1])
    for e in range(bd.shape[0]):
        s, t = int(indptr[e]), int(indptr[e + 1])
        out[s:t] = ad[s:t].float() @ bd[e].float().t()
    return out


def test_grouped_mm_fp8_per_tensor():
    torch.manual_seed(0)
    G, n, k = 3, 24, 64
    indptr = torch.tensor([0, 5, 5, 21], dtype=torch.int32)
    a = (torch.randn(21, k) * 0.5).to(torch.float8_e4m3fn)
    b = (torch.randn(G, n, k) * 0.5).to(torch.float8_e4m3fn)
    alpha = torch.tensor([0.37])
    want = _loop(a, b, indptr) * 0.37
    got = fi.grouped_mm.grouped_mm_fp8(a, b, indptr, alpha=alpha, out_dtype=torch.float32)
    torch.testing.assert_close(got, want, atol=2e-2, rtol=1e-2)          # the grouped GEMM stores bf16
    out = torch.empty(21, n, dtype=torch.bfloat16)
    assert fi.grouped_mm.grouped_mm_fp8(a, b, indptr, out=out, backend="cudnn", tactic=0) is out
    torch.testing.assert_close(out.float(), _loop(a, b, indptr), atol=5e-2, rtol=2e-2)
    assert fi.grouped_mm.core.grouped_mm_fp8 is fi.grouped_mm.grouped_mm_fp8
    torch.testing.assert_close(fi.grouped_mm.grouped_mm_bf16(a.bfloat16(), b.bfloat16(), indptr, backend="cudnn", tactic=-1).float(), _loop(a, b, indptr),
                               atol=5e-2, rtol=2e-2)


@pytest.mark.parametrize("block_size", [16, 32])
def test_grouped_mm_fp4_swizzled_scales(block_size):
    torch.manual_seed(1)
    G, n, k = 2, 32, 128
    indptr = torch.tensor([0, 40, 150], dtype=torch.int32)           # 150 rows: more than one 128-row swizzle tile
    a = torch.randn(150, k).bfloat16()
    b = torch.randn(G, n, k).bfloat16()
    ue8 = block_size == 32
    gs = None if ue8 else torch.tensor([1.0])
    aq, asf = fp4_quantize(a, gs, block_size, ue8, True)
    bq, bsf = zip(*(fp4_quantize(b[g], gs, block_size, ue8, True) for g in range(G)))
    bq, bsf = torch.stack(bq), torch.stack([s.reshape(-1) for s in bsf])
    # oracle: the linear-scale quantisation of the same data, de-quantised by hand
    from flashinfer_b200.gemm.grouped import _dq_fp4

    aql, asl = fp4_quantize(a, gs, block_size, ue8, False)
    assert torch.equal(aql, aq)
    ad = _dq_fp4(aql, asl.view(150, k // block_size), block_size, "ue8m0" if ue8 else "ue4m3")
    bd = torch.stack([_dq_fp4(*(lambda q, s: (q, s.view(n, k // block_size)))(*fp4_quantize(b[g], gs, block_size, ue8, False)), block_size,
                              "ue8m0" if ue8 else "ue4m3") for g in range(G)])
    assert torch.nn.functional.cosine_similarity(ad.float().flatten(), a.float().flatten(), dim=0) > 0.98      # the oracle is a quantisation of a
    alpha = torch.tensor([2.0])
    got = fi.grouped_mm.grouped_mm_fp4(aq, bq, asf, bsf, indptr, alpha=alpha, out_dtype=torch.float32, block_size=block_size)
    torch.testing.assert_close(got, 2.0 * _loop(ad, bd, indptr), atol=0.3, rtol=1e-2)
    with pytest.raises(ValueError):
        fi.grouped_mm.grouped_mm_fp4(aq, bq, asf, bsf, indptr, block_size=8)
    with pytest.raises(ValueError):
        fi.grouped_mm.grouped_mm_fp4(aq, bq, asf.reshape(-1)[:100], bsf, indptr, block_size=block_size)


def test_cutlass_fused_moe_input_sf_layouts_and_small_parity_arguments(monkeypatch):
    """NVFP4 activations: ``input_sf`` is 128x4-swizzled by default (reference core.py :916), linear with ``swizzled_input_sf=False`` - both
    must give the same result; plus ``mxfp8_dequantize_host(sf_swizzle_layout=)`` and ``MoEInputs.from_list(lst=)``."""
    from flashinfer_b200.fused_moe import core
    from flashinfer_b200.quantization.fp4 import SfLayout
    from flashinfer_b200.quantization.fp8 import mxfp8_dequantize_host, mxfp8_quantize

    torch.manual_seed(2)
    T, H, I, E, K = 5, 64, 32, 4, 2
    x = torch.randn(T, H).bfloat16()
    one = torch.tensor([1.0])
    xq_s, sf_s = fp4_quantize(x, one, 16, False, True)
    xq_l, sf_l = fp4_quantize(x, one, 16, False, False)
    assert torch.equal(xq_s, xq_l)
    w1, w2 = (torch.randn(E, 2 * I, H) * 0.1).bfloat16(), (torch.randn(E, H, I) * 0.1).bfloat16()
    q1 = [fp4_quantize(w1[e], one, 16, False, False) for e in range(E)]
    q2 = [fp4_quantize(w2[e], one, 16, False, False) for e in range(E)]
    w1q, w1sf = torch.stack([a for a, _ in q1]), torch.stack([b.view(2 * I, H // 16) for _, b in q1])
    w2q, w2sf = torch.stack([a for a, _ in q2]), torch.stack([b.view(H, I // 16) for _, b in q2])
    ids = torch.randint(0, E, (T, K), dtype=torch.int32)
    wts = torch.rand(T, K)
    ones = torch.ones(E)
    scales = [one, w1sf, ones, one, w2sf, ones]
    # the NVFP4 pipeline itself is CUDA-only: stand in for it and look at the activations it is handed
    monkeypatch.setattr(core, "moe_forward_nvfp4", lambda xb, *a, **k: xb.clone())
    a = core.cutlass_fused_moe(xq_s, ids, wts, w1q, w2q, torch.bfloat16, quant_scales=scales, input_sf=sf_s)
    b = core.cutlass_fused_moe(xq_l, ids, wts, w1q, w2q, torch.bfloat16, quant_scales=scales, input_sf=sf_l, swizzled_input_sf=False)
    a, b = (t[0] if isinstance(t, (list, tuple)) else t for t in (a, b))
    torch.testing.assert_close(a, b)
    assert torch.nn.functional.cosine_similarity(a.float().flatten(), x.float().flatten(), dim=0) > 0.98           # = the de-quantised input
    with pytest.raises(ValueError):
        core.cutlass_fused_moe(xq_s, ids, wts, w1q, w2q, torch.bfloat16, quant_scales=scales)
    y = torch.randn(130, 64).bfloat16()
    yq, ysf = mxfp8_quantize(y, True)
    d0 = mxfp8_dequantize_host(yq, ysf)
    torch.testing.assert_close(mxfp8_dequantize_host(yq, ysf, False, sf_swizzle_layout=SfLayout.layout_128x4), d0)
    yq2, ysf2 = mxfp8_quantize(y, False)
    torch.testing.assert_close(mxfp8_dequantize_host(yq2, ysf2, True, sf_swizzle_layout=SfLayout.layout_linear), d0)
    with pytest.raises(ValueError):
        mxfp8_dequantize_host(yq, ysf, sf_swizzle_layout=SfLayout.layout_8x4)
    fields = core.MoEInputs._FIELDS
    assert core.MoEInputs.from_list(lst=list(range(len(fields)))).__dict__[fields[1]] == 1
