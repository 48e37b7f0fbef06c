"""fused_moe.moe_utils: the standalone MoE building blocks (reference flashinfer/fused_moe/cute_dsl/moe_utils.py,
tests/moe/test_cute_dsl_fused_moe.py helper checks).  Composed here into a whole expert FFN and compared with the fused entry point."""
import pytest
import torch

from flashinfer_b200.fused_moe import moe_utils as mu
from flashinfer_b200.fused_moe.core import moe_reference


@pytest.mark.parametrize("tokens,top_k,experts,tile", [(1, 8, 256, 128), (5, 2, 4, 8), (64, 8, 32, 16), (300, 4, 7, 128), (3, 1, 2, 128)])
def test_max_tiles_bound_is_tight_and_sufficient(tokens, top_k, experts, tile):
    rows = tokens * top_k
    worst = [1] * min(experts - 1, max(rows - 1, 0))
    worst.append(rows - sum(worst))
    need = sum(-(-c // tile) for c in worst if c > 0)
    assert mu.get_max_num_tiles(tokens, top_k, experts, tile) == need
    assert mu.get_max_num_permuted_tokens(tokens, top_k, experts, tile) == need * tile


def _check_sort(ids, num_experts, offset, local, tile):
    t, k = ids.shape
    te, lim, e2p, p2e, total, ntiles = mu.moe_sort(ids, torch.ones(t, k), num_experts, k, offset, local, tile)
    assert te.dtype == lim.dtype == e2p.dtype == p2e.dtype == torch.int32
    n_tiles, n_rows = int(ntiles), int(total)
    assert n_rows == n_tiles * tile and p2e.numel() == mu.get_max_num_permuted_tokens(t, k, local, tile)
    loc = ids.long() - offset
    for i in range(t):
        for j in range(k):
            p = int(e2p[i, j])
            if 0 <= loc[i, j] < local:
                assert 0 <= p < n_rows and int(p2e[p]) == i * k + j                      # the two maps are inverse
                tl = p // tile
                assert int(te[tl]) == int(loc[i, j]) and p < int(lim[tl])                  # the row sits in a tile of its expert, below the limit
            else:
                assert p == -1
    live = p2e[:n_rows] >= 0
    assert int(live.sum()) == int(((loc >= 0) & (loc < local)).sum()) and bool((p2e[n_rows:] == -1).all())
    for tl in range(n_tiles):                                                              # valid rows of a tile = a prefix, padding after it
        rows = live[tl * tile:(tl + 1) * tile]
        nv = int(lim[tl]) - tl * tile
        assert 0 < nv <= tile and bool(rows[:nv].all()) and not bool(rows[nv:].any())
    assert bool((te[:n_tiles][1:] >= te[:n_tiles][:-1]).all())                              # experts in ascending order
    for e in range(local):                                                                  # token order preserved inside an expert
        rows = p2e[:n_rows][(te.long().repeat_interleave(tile)[:n_rows] == e) & live]
        assert bool((rows[1:] > rows[:-1]).all())
    return te, lim, e2p, p2e


def test_moe_sort_layout():
    g = torch.Generator().manual_seed(0)
    _check_sort(torch.randint(0, 8, (37, 2), generator=g).int(), 8, 0, 8, 16)
    _check_sort(torch.randint(0, 16, (50, 4), generator=g).int(), 16, 4, 6, 8)             # expert parallel: experts 4..9 are local
    _check_sort(torch.full((9, 1), 3).int(), 8, 0, 8, 4)                                    # everything on one expert
    _check_sort(torch.randint(12, 16, (6, 2), generator=g).int(), 16, 0, 8, 8)              # nothing local: zero tiles
    ids = torch.randint(0, 8, (21, 2), generator=g).int()
    bufs = mu.allocate_moe_sort_buffers(21, 8, 2, 8, 16, device="cpu")
    res = mu.moe_sort(ids, torch.ones(21, 2), 8, 2, tile_tokens_dim=16, **bufs)
    assert all(r.data_ptr() == b.data_ptr() for r, b in zip(res, bufs.values()))            # pre-allocated buffers are filled in place
    with pytest.raises(ValueError):
        mu.moe_sort(ids, torch.ones(21, 2), 8, 3)


@pytest.mark.parametrize("offset,local", [(0, 8), (2, 4)])
def test_building_blocks_compose_into_the_fused_moe(offset, local):
    g = torch.Generator().manual_seed(1)
    t, k, e, h, inter, tile = 33, 2, 8, 32, 16, 8
    x = torch.randn(t, h, generator=g).to(torch.bfloat16)
    w1 = (torch.randn(local, 2 * inter, h, generator=g) / h ** 0.5).to(torch.bfloat16)
    w2 = (torch.randn(local, h, inter, generator=g) / inter ** 0.5).to(torch.bfloat16)
    scales, ids = torch.topk(torch.softmax(torch.randn(t, e, generator=g), -1), k)
    ids = ids.int()
    te, lim, e2p, p2e, total, ntiles = mu.moe_sort(ids, scales, e, k, offset, local, tile)
    rows = mu.get_max_num_permuted_tokens(t, k, local, tile)
    xp = torch.full((rows, h), float("nan"), dtype=torch.bfloat16)
    mu.moe_permute(x, xp, lim, p2e, ntiles, rows, k, tile)
    assert not torch.isnan(xp.float()).any()
    h1 = torch.zeros(rows, 2 * inter, dtype=torch.bfloat16)
    for tl in range(int(ntiles)):                                                            # grouped GEMM 1, tile by tile
        h1[tl * tile:(tl + 1) * tile] = (xp[tl * tile:(tl + 1) * tile].float() @ w1[int(te[tl])].float().t()).to(torch.bfloat16)
    a = torch.empty(rows, inter, dtype=torch.bfloat16)
    mu.moe_swiglu(h1, a, lim, ntiles, rows, tile)
    h2 = torch.zeros(rows, h, dtype=torch.bfloat16)
    for tl in range(int(ntiles)):
        h2[tl * tile:(tl + 1) * tile] = (a[tl * tile:(tl + 1) * tile].float() @ w2[int(te[tl])].float().t()).to(torch.bfloat16)
    out = torch.full((t, h), 7.0, dtype=torch.bfloat16)
    mu.moe_output_memset(out, lim, e2p, p2e, ntiles, rows, k, tile)
    touched = ((ids >= offset) & (ids < offset + local)).any(-1)
    assert bool((out[touched] == 0).all()) and bool((out[~touched] == 7.0).all())
    mu.moe_unpermute(h2, out, e2p, scales, t, k)
    ref = moe_reference(x, ids, scales, w1, w2, "silu", offset)
    torch.testing.assert_close(out.float(), ref, atol=3e-2, rtol=3e-2)


def test_activations_and_padding_rows():
    g = torch.Generator().manual_seed(2)
    tile, rows, inter = 4, 12, 8
    lim = torch.tensor([3, 8, 0], dtype=torch.int32)                                         # tile 0: 3 rows, tile 1: full, tile 2: unused
    x2 = torch.randn(rows, 2 * inter, generator=g).to(torch.bfloat16)
    x1 = x2[:, :inter].contiguous()
    keep = torch.tensor([1, 1, 1, 0, 1, 1, 1, 1, 0, 0, 0, 0], dtype=torch.bool)
    n = torch.tensor([2], dtype=torch.int32)
    cases = [(mu.moe_swiglu, x2, lambda v: torch.nn.functional.silu(v[:, inter:]) * v[:, :inter]),
             (mu.moe_geglu, x2, lambda v: torch.nn.functional.gelu(v[:, inter:]) * v[:, :inter]),
             (mu.moe_gelu, x1, torch.nn.functional.gelu), (mu.moe_silu, x1, torch.nn.functional.silu), (mu.moe_relu, x1, torch.relu)]
    for fn, src, ref in cases:
        out = torch.full((rows, inter), float("nan"), dtype=torch.bfloat16)
        fn(src, out, lim, n, rows, tile)
        want = torch.where(keep[:, None], ref(src.float()), torch.zeros(()))
        torch.testing.assert_close(out.float(), want, atol=2e-2, rtol=2e-2)
    out = torch.empty(rows, inter, dtype=torch.bfloat16)
    mu.moe_activation(x1, out, lim, n, mu.MoeActivationType.Identity, rows, tile)
    assert torch.equal(out[keep], x1[keep])
    with pytest.raises(ValueError):
        mu.moe_swiglu(x1, out, lim, n, rows, tile)
    z = torch.ones(4, 4)
    mu.moe_output_memset_inplace(z)
    assert float(z.abs().sum()) == 0.0


def test_permute_nvfp4_rows_with_scale_factors():
    from flashinfer_b200.quantization.fp4 import block_scale_interleave

    g = torch.Generator().manual_seed(3)
    t, k, hidden, tile = 6, 2, 64, 4
    ids = torch.randint(0, 4, (t, k), generator=g).int()
    te, lim, e2p, p2e, total, ntiles = mu.moe_sort(ids, torch.ones(t, k), 4, k, tile_tokens_dim=tile)
    rows = mu.get_max_num_permuted_tokens(t, k, 4, tile)
    packed = torch.randint(0, 256, (t, hidden // 2), generator=g, dtype=torch.uint8)
    sf = torch.randint(1, 120, (t, hidden // 16), generator=g, dtype=torch.uint8)
    outp = torch.empty(rows, hidden // 2, dtype=torch.uint8)
    outsf = torch.zeros(((rows + 127) // 128) * 128 * 4, dtype=torch.uint8)
    mu.moe_permute(packed, outp, lim, p2e, ntiles, rows, k, tile, input_sf=sf, permuted_sf=outsf)
    tok = (p2e.long().clamp(min=0) // k)
    has = (p2e >= 0)
    assert torch.equal(outp[has], packed[tok[has]]) and float(outp[~has].float().abs().sum()) == 0.0
    lin = torch.where(has[:, None], sf[tok], torch.zeros_like(sf[tok]))
    assert torch.equal(outsf, block_scale_interleave(lin).reshape(-1))
    r, c = int(torch.nonzero(has)[-1]), 3                 # spot check against the 128x4 swizzle formula
    assert int(outsf[(r % 32) * 16 + (r // 32) * 4 + c]) == int(lin[r, c])
    with pytest.raises(ValueError):
        mu.moe_permute(packed, outp, lim, p2e, ntiles, rows, k, tile, input_sf=sf, permuted_sf=torch.zeros(rows * 4, dtype=torch.uint8))
    with pytest.raises(ValueError):
        mu.moe_permute(packed, outp, lim, p2e, ntiles, rows, k, tile, input_sf=sf)
