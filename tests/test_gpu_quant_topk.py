"""GPU tests: quantisation kernels (vs. the CPU reference implementation) and radix top-k."""
import numpy as np
import pytest
import torch

from flashinfer_b200 import quantization as Q
from flashinfer_b200 import topk as TK

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(130, 64), (512, 7168), (1, 128)])
@pytest.mark.parametrize("swizzled", [True, False])
def test_nvfp4_quantize_matches_cpu_and_roundtrips(shape, swizzled):
    x = torch.randn(shape, dtype=torch.bfloat16)
    gs = (448.0 * 6.0 / x.float().abs().max()).reshape(1)
    q_c, sf_c = Q.fp4_quantize(x, gs, is_sf_swizzled_layout=swizzled)
    q_g, sf_g = Q.fp4_quantize(x.cuda(), gs.cuda(), is_sf_swizzled_layout=swizzled)
    assert (sf_g.cpu() == sf_c).float().mean() > 0.999
    assert (q_g.cpu() == q_c).float().mean() > 0.99  # rare 1-ulp differences in the scaled value
    d = Q.e2m1_and_ufp8sf_scale_to_float(q_g, sf_g, gs.cuda(), is_sf_swizzled_layout=swizzled)
    cos = torch.nn.functional.cosine_similarity(d.flatten().cpu(), x.float().flatten(), dim=0)
    assert cos > 0.99


def test_mxfp4_mxfp8_batched_interleave():
    x = torch.randn(300, 256, dtype=torch.bfloat16, device="cuda")
    q, sf = Q.mxfp4_quantize(x)
    d = Q.mxfp4_dequantize(q, sf)
    assert torch.nn.functional.cosine_similarity(d.flatten(), x.float().flatten(), dim=0) > 0.99
    q8, s8 = Q.mxfp8_quantize(x)
    q8c, s8c = Q.mxfp8_quantize(x.cpu())
    assert torch.equal(s8.cpu(), s8c)
    assert (q8.cpu().float() == q8c.float()).float().mean() > 0.999
    d8 = Q.mxfp8_dequantize_host(q8, s8)
    torch.testing.assert_close(d8, x.float(), rtol=0.07, atol=0.02)
    xb = torch.randn(3, 128, 64, dtype=torch.bfloat16, device="cuda")
    gs = torch.tensor([100.0], device="cuda")
    qb, sb = Q.nvfp4_batched_quantize(xb, gs)
    for i in range(3):
        qi, si = Q.fp4_quantize(xb[i], gs)
        assert torch.equal(qb[i], qi) and torch.equal(sb[i], si.reshape(-1))
    lin = torch.randint(0, 255, (256, 8), dtype=torch.uint8, device="cuda")
    assert torch.equal(Q.block_scale_interleave(lin).cpu(), Q.block_scale_interleave(lin.cpu()))


def test_packbits():
    b = torch.rand(1003, device="cuda") > 0.5
    assert (Q.packbits(b).cpu().numpy() == np.packbits(b.cpu().numpy())).all()
    assert (Q.packbits(b, "little").cpu().numpy() == np.packbits(b.cpu().numpy(), bitorder="little")).all()
    indptr = torch.tensor([0, 10, 10, 333, 1003], dtype=torch.int32, device="cuda")
    out, new_indptr = Q.segment_packbits(b, indptr)
    ref, ref_indptr = Q.segment_packbits(b.cpu(), indptr.cpu())
    assert torch.equal(out.cpu(), ref) and torch.equal(new_indptr.cpu(), ref_indptr)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n,k", [(1000, 1), (32000, 50), (128256, 2048), (100, 100)])
def test_top_k(dtype, n, k):
    x = torch.randn(5, n, device="cuda", dtype=dtype)
    v, i = TK.top_k(x, k, sorted=True)
    vr, ir = torch.topk(x, k, dim=-1, sorted=True)
    torch.testing.assert_close(v.float(), vr.float())
    torch.testing.assert_close(x.gather(-1, i).float(), vr.float())
    assert all(len(set(r.tolist())) == k for r in i)


def test_top_k_transforms_and_ties():
    rows, max_len, k = 6, 4096, 64
    x = torch.randn(rows, max_len, device="cuda")
    lengths = torch.tensor([4096, 100, 64, 10, 2000, 4096], dtype=torch.int32, device="cuda")
    table = torch.randint(0, 100000, (rows, max_len), dtype=torch.int32, device="cuda")
    out = TK.top_k_page_table_transform(x, table, lengths, k)
    ref = TK.top_k_page_table_transform(x.cpu(), table.cpu(), lengths.cpu(), k)
    for r in range(rows):
        assert sorted(out[r].tolist()) == sorted(ref[r].tolist())
    offs = torch.arange(rows, dtype=torch.int32, device="cuda") * 10000
    out = TK.top_k_ragged_transform(x, offs, lengths, k)
    ref = TK.top_k_ragged_transform(x.cpu(), offs.cpu(), lengths.cpu(), k)
    for r in range(rows):
        assert sorted(out[r].tolist()) == sorted(ref[r].tolist())
    # ties: all-equal row -> SMALL takes the first k indices, LARGE the last k
    t = torch.zeros(2, 500, device="cuda")
    _, i = TK.top_k(t, 7, tie_break=TK.TopKTieBreak.SMALL)
    assert sorted(i[0].tolist()) == list(range(7))
    _, i = TK.top_k(t, 7, tie_break=TK.TopKTieBreak.LARGE)
    assert sorted(i[0].tolist()) == list(range(493, 500))


@pytest.mark.parametrize("rows,n,k,cs", [(1, 131072, 2048, 8), (3, 65536, 2048, 4), (2, 40000, 512, 2), (5, 9000, 64, 8), (1, 450000, 2048, 8),
                                         (4, 131072, 2048, 0)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_top_k_cluster_kernel(rows, n, k, cs, dtype):
    """One row per thread-block cluster (csrc/elementwise/topk.cu: topk_cluster_kernel): same selection as the one-CTA kernel and
    torch.topk, for slices that fit the shared-memory key cache and for slices that do not (n = 450000 over 8 CTAs = 56 K keys per slice)."""
    torch.manual_seed(rows * 1000 + k)
    x = torch.randn(rows, n, device="cuda").to(dtype)
    vals, idx = TK._run(x, k, 0, want_values=True, clusters=cs)
    v1, i1 = TK._run(x, k, 0, want_values=True, clusters=1)
    assert torch.equal(idx, i1) and torch.equal(vals, v1)  # deterministic, index-ordered: bit-identical to the single-CTA kernel
    vr = torch.topk(x.float(), k, dim=-1).values
    assert torch.equal(vals.float().sort(-1, descending=True).values, vr)
    assert torch.equal(x.gather(1, idx.long()), vals)


def test_top_k_cluster_transforms_ties_lengths():
    torch.manual_seed(1)
    rows, n, k = 3, 50000, 1024
    x = torch.randn(rows, n, device="cuda")
    x[:, ::3] = 0.25  # heavy ties around the threshold
    lengths = torch.tensor([50000, 20001, 700], dtype=torch.int32, device="cuda")
    table = torch.randint(0, 1 << 20, (rows, n), dtype=torch.int32, device="cuda")
    offs = torch.tensor([5, 100000, 7], dtype=torch.int32, device="cuda")
    for tb in (TK.TopKTieBreak.NONE, TK.TopKTieBreak.SMALL, TK.TopKTieBreak.LARGE):
        a = TK._run(x, k, 1, lengths=lengths, page_table=table, tie_break=tb, clusters=8)[1]
        b = TK._run(x, k, 1, lengths=lengths, page_table=table, tie_break=tb, clusters=1)[1]
        assert torch.equal(a, b)
        a = TK._run(x, k, 2, lengths=lengths, ragged_offsets=offs, tie_break=tb, clusters=4)[1]
        b = TK._run(x, k, 2, lengths=lengths, ragged_offsets=offs, tie_break=tb, clusters=1)[1]
        assert torch.equal(a, b)
    assert (a[2, 700:] == -1).all()
    idx = TK.topk_clusters_exact(x, k)
    assert idx.dtype == torch.int32 and torch.equal(idx, TK._run(x, k, 0, clusters=1)[1])
    assert torch.equal(TK.topk_clusters_page_table_transform(x, lengths, table, k), TK.top_k_page_table_transform(x, table, lengths, k))
    assert torch.equal(TK.topk_clusters_ragged_transform(x, lengths, offs, k), TK.top_k_ragged_transform(x, offs, lengths, k))


@pytest.mark.parametrize("shape", [(8, 64), (37, 256), (200, 4096)])
def test_nvfp4_quantize_8x4_layout(shape):
    """SfLayout.layout_8x4 (tiles of 8 rows x 4 scale columns): same codes and scale bytes as the linear layout, at the 8x4 offsets."""
    from flashinfer_b200.quantization.fp4 import SfLayout, _index_8x4, nvfp4_quantize

    torch.manual_seed(shape[0])
    x = torch.randn(*shape, device="cuda").bfloat16()
    gs = torch.tensor([448.0 * 6.0 / float(x.float().abs().max())], device="cuda")
    q8, sf8 = nvfp4_quantize(x, gs, sfLayout=SfLayout.layout_8x4)
    ql, sfl = nvfp4_quantize(x, gs, sfLayout=SfLayout.layout_linear)
    assert torch.equal(q8, ql)
    m, kc = shape[0], shape[1] // 16
    assert sf8.numel() == (m + 7) // 8 * 8 * ((kc + 3) // 4 * 4)
    assert torch.equal(sf8.reshape(-1)[_index_8x4(m, kc).to("cuda")], sfl.reshape(-1))
    qc, sfc = nvfp4_quantize(x.cpu(), gs.cpu(), sfLayout=SfLayout.layout_8x4)
    assert (sfc.reshape(-1) == sf8.reshape(-1).cpu()).float().mean() > 0.95  # 1-ulp differences of the e4m3 scale (CPU oracle)
    assert (sfc.reshape(-1).int() - sf8.reshape(-1).cpu().int()).abs().max() <= 1
