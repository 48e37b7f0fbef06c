"""Property-based tests (hypothesis) of host-side logic whose bugs show up only at odd sizes: the C++ decode planner, the MoE tile
sort, bit packing of custom masks, scale-factor swizzles, context-parallel shard construction."""
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

import flashinfer_b200 as fi
from test_moe_utils_cpu import _check_sort
from test_planner_cpu import _plan, _tiles

_FAST = settings(max_examples=40, deadline=None)


@_FAST
@given(kv_lens=st.lists(st.integers(0, 5000), min_size=1, max_size=40), ps=st.sampled_from([1, 8, 16, 64, 128, 200]),
       hkv=st.sampled_from([1, 2, 8]), num_ctas=st.sampled_from([4, 37, 148]))
def test_decode_planner_partitions_the_work(kv_lens, ps, hkv, num_ctas):
    seg, cta, mrg, counts = _plan(kv_lens, hkv, 4, ps, num_ctas)
    seen = set()
    for s in seg.tolist():
        b, h, t0, t1 = s[0], s[1], s[2], s[3]
        assert 0 <= b < len(kv_lens) and 0 <= h < hkv and t1 > t0
        for t in range(t0, t1):
            assert (b, h, t) not in seen
            seen.add((b, h, t))
    assert len(seen) == sum(_tiles(n, ps) for n in kv_lens) * hkv                       # every (request, head, tile) exactly once
    assert int(cta[0]) == 0 and int(cta[num_ctas]) == len(seg) and bool((cta[1:] >= cta[:-1]).all())
    slots = {s[4]: (s[0], s[1]) for s in seg.tolist() if s[4] >= 0}
    for m in mrg.tolist():                                                               # a merge reads consecutive partial slots of one (request, head)
        assert all(slots[m[0] + i] == (m[5], m[4]) for i in range(m[1]))


@_FAST
@given(tokens=st.integers(1, 70), top_k=st.integers(1, 4), experts=st.integers(1, 12), tile=st.sampled_from([4, 8, 128]),
       offset=st.integers(0, 3), seed=st.integers(0, 10 ** 6), data=st.data())
def test_moe_sort_invariants(tokens, top_k, experts, tile, offset, seed, data):
    local = data.draw(st.integers(1, experts))
    offset = min(offset, experts - local)
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, experts, (tokens, top_k), generator=g).int()
    _check_sort(ids, experts, offset, local, tile)


@_FAST
@given(lens=st.lists(st.integers(0, 300), min_size=1, max_size=8), seed=st.integers(0, 10 ** 6))
def test_segment_packbits_round_trip(lens, seed):
    from flashinfer_b200.prefill import _unpack_segmented

    g = torch.Generator().manual_seed(seed)
    flat = torch.rand(sum(lens), generator=g) > 0.5
    indptr = torch.tensor([0] + lens).cumsum(0).int()
    packed, new_indptr = fi.segment_packbits(flat, indptr, bitorder="little")
    assert new_indptr.tolist() == [0] + torch.tensor([(n + 7) // 8 for n in lens]).cumsum(0).tolist()
    assert torch.equal(_unpack_segmented(packed, lens), flat)


@_FAST
@given(m=st.integers(1, 300), kc=st.integers(1, 20), seed=st.integers(0, 10 ** 6))
def test_scale_factor_swizzle_is_a_permutation_into_padded_tiles(m, kc, seed):
    from flashinfer_b200.quantization.fp4 import block_scale_interleave

    g = torch.Generator().manual_seed(seed)
    sf = torch.randint(1, 255, (m, kc), generator=g, dtype=torch.uint8)
    swz = block_scale_interleave(sf).reshape(-1)
    rows, cols = (m + 127) // 128 * 128, (kc + 3) // 4 * 4
    assert swz.numel() == rows * cols
    r, c = torch.meshgrid(torch.arange(m), torch.arange(kc), indexing="ij")
    off = ((r // 128) * (cols // 4) + c // 4) * 512 + (r % 32) * 16 + ((r % 128) // 32) * 4 + c % 4    # the 128x4 swizzle
    assert torch.equal(swz[off.reshape(-1)], sf.reshape(-1))
    assert int((swz != 0).sum()) == m * kc                                               # everything else is padding


@_FAST
@given(seqs=st.lists(st.integers(1, 50), min_size=1, max_size=6), world=st.integers(1, 5))
def test_varlen_shards_reassemble_every_sequence(seqs, world):
    from flashinfer_b200.parallel_attention import split_varlen_input

    total = sum(seqs)
    x = torch.arange(1, total + 1, dtype=torch.float32).view(1, total, 1)               # HND, token id = value (0 marks padding)
    shards = [split_varlen_input(x, seqs, world, r, "HND") for r in range(world)]
    per = [-(-n // world) for n in seqs]
    assert all(s.shape[1] == sum(per) for s in shards)                                   # equal shapes: zero padding at the end of a shard
    chunk = lambda n, p, r: max(min(n - p * r, p if r < world - 1 else n), 0)            # noqa: E731  tokens of a sequence on rank r
    cursor = [0] * world
    offset = 0
    for n, p in zip(seqs, per):
        got = []
        for r in range(world):
            c = chunk(n, p, r)
            got += shards[r][0, cursor[r]:cursor[r] + c, 0].tolist()
            cursor[r] += c
        assert got == list(range(offset + 1, offset + n + 1))                            # chunks in rank order rebuild the sequence
        offset += n
    for r in range(world):
        assert float(shards[r][0, cursor[r]:, 0].abs().sum()) == 0.0                     # what is left is padding
