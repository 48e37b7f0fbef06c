"""Context-parallel attention on gloo ranks against single-process attention (reference tests/attention/test_parallel_attention.py:
every rank checks its shard of the global result; plain / uneven / Ulysses-varlen / Ring-varlen cases, HND and NHD, plus what this
library adds: causal single-dimension runs and GQA with fused q / k / v exchange)."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _sdpa(q, k, v, causal=False):
    """q [Hq, S, D], k / v [Hkv, Skv, D] (HND) -> [Hq, S, D] in fp32; causal aligns the ends."""
    g = q.shape[0] // k.shape[0]
    kf, vf = k.float().repeat_interleave(g, 0), v.float().repeat_interleave(g, 0)
    lg = torch.einsum("hqd,hkd->hqk", q.float(), kf) / math.sqrt(q.shape[-1])
    if causal:
        sq, sk = q.shape[1], k.shape[1]
        lg = lg.masked_fill(torch.arange(sk)[None, :] > torch.arange(sq)[:, None] + (sk - sq), float("-inf"))
    return torch.einsum("hqk,hkd->hqd", torch.softmax(lg, -1), vf)


def _to_layout(t, layout):
    return t.transpose(0, 1).contiguous() if layout == "NHD" else t.contiguous()


def _from_layout(t, layout):
    return t.transpose(0, 1) if layout == "NHD" else t


def _groups(world, ulysses, ring):
    from flashinfer_b200.parallel_attention import get_parallel_groups

    return get_parallel_groups(ulysses_size=ulysses, ring_size=ring, device_type="cpu")


def _cases_world2(rank, world):
    from flashinfer_b200.parallel_attention import (ParallelAttention, UnevenCPConfig, VarlenCPConfig, ring_varlen_config, split_varlen_input,
                                                    ulysses_varlen_config, uneven_cp_config)

    errs = {}
    g = torch.Generator().manual_seed(0)
    H, S, D = 4, 48, 32
    q, k, v = (torch.randn(H, S, D, generator=g) for _ in range(3))
    kg, vg = (torch.randn(2, S, D, generator=g) for _ in range(2))          # GQA: 2 kv heads
    shard = lambda t: t.chunk(world, dim=1)[rank]  # noqa: E731
    ring_g, _ = _groups(world, 1, world)
    _, uly_g = _groups(world, world, 1)
    for mode in ("ulysses", "ulysses_fused", "ring"):
        for layout in ("HND", "NHD"):
            for causal in (False, True):
                pa = ParallelAttention("sm100", uly_g if mode != "ring" else None, ring_g if mode == "ring" else None, fuse_qkv=mode == "ulysses_fused")
                out = pa.run(*(_to_layout(shard(t), layout) for t in (q, k, v)), layout, is_causal=causal)
                ref = shard(_sdpa(q, k, v, causal))
                errs[f"{mode}/{layout}/causal={causal}"] = float((_from_layout(out, layout).float() - ref).abs().max())
    # GQA + fused exchange falls back to three collectives (shapes differ)
    pa = ParallelAttention("cutlass", uly_g, None, fuse_qkv=True)
    out = pa.run(shard(q), shard(kg), shard(vg), "HND")
    errs["gqa_fused"] = float((out.float() - shard(_sdpa(q, kg, vg))).abs().max())

    # ---- uneven: the last rank's shard ends with (world - 1 + 2) padded rows
    pad = world + 1
    real = S - pad
    for mode in ("ulysses", "ring"):
        for layout in ("HND", "NHD"):
            ug, rg = (uly_g, None) if mode == "ulysses" else (None, ring_g)
            mine = S // world - (pad if rank == world - 1 else 0)
            lens = uneven_cp_config(seq_len=real, seq_len_padded=S, seq_len_cur_rank=mine, ulysses_group=ug, ring_group=rg)
            assert (lens is None) == (mode == "ulysses")
            cfg = UnevenCPConfig(seq_len=real, seq_len_padded=S, seq_len_cur_ring_group=lens)
            junk = [t.clone() for t in (q, k, v)]
            for t in junk:
                t[:, real:] = 1e3                                             # garbage in the padding must not leak into the result
            out = ParallelAttention("flash-attn3", ug, rg, uneven_cp_config=cfg).run(*(_to_layout(shard(t), layout) for t in junk), layout)
            out = _from_layout(out, layout).float()
            ref = torch.zeros(H, S, D)
            ref[:, :real] = _sdpa(q[:, :real], k[:, :real], v[:, :real])
            errs[f"uneven/{mode}/{layout}"] = float((out - shard(ref)).abs().max())     # includes: padded rows are zeros

    # ---- varlen, Ulysses only: the packed batch is exchanged as a whole
    seqs = [13, 21, 9]                                                        # 43 tokens, padded to 44
    total = sum(seqs)
    padded = -(-total // world) * world
    qv, kv_, vv = (torch.randn(H, padded, D, generator=g) for _ in range(3))
    cq, ck, mq, mk = ulysses_varlen_config(seqs, seqs)
    cfg = VarlenCPConfig()
    cfg.set_varlen_cp_config(cq, ck, mq, mk, uly_g, None)
    assert cfg.cu_seqlens_q_cur_ulysses_group is cq and cfg.cu_seqlens_q_cur_ring_group is None
    ref = torch.zeros(H, padded, D)
    for a, b in zip(cq.tolist()[:-1], cq.tolist()[1:]):
        ref[:, a:b] = _sdpa(qv[:, a:b], kv_[:, a:b], vv[:, a:b])
    for layout in ("HND", "NHD"):
        out = ParallelAttention("sm100", uly_g, None, varlen_cp_config=cfg).run(*(_to_layout(shard(t), layout) for t in (qv, kv_, vv)), layout)
        errs[f"varlen_ulysses/{layout}"] = float((_from_layout(out, layout).float() - shard(ref)).abs().max())

    # ---- varlen, Ring only: rank r holds chunk r of every sequence
    seqs = [11, 16, 7, 2]
    total = sum(seqs)
    qv, kv_, vv = (torch.randn(H, total, D, generator=g) for _ in range(3))
    cq, ck, mq, mk = ring_varlen_config(seqs, seqs, ring_g)
    assert tuple(cq.shape) == (world, len(seqs) + 1) and mq == max(-(-n // world) for n in seqs)
    cfg = VarlenCPConfig()
    cfg.set_varlen_cp_config(cq, ck, mq, mk, None, ring_g)
    full = torch.zeros(H, total, D)
    o = 0
    for n in seqs:
        full[:, o:o + n] = _sdpa(qv[:, o:o + n], kv_[:, o:o + n], vv[:, o:o + n])
        o += n
    for layout in ("HND", "NHD"):
        mk_shard = lambda t: split_varlen_input(_to_layout(t, layout), seqs, world, rank, layout)  # noqa: E731
        out = ParallelAttention("sm100", None, ring_g, varlen_cp_config=cfg).run(mk_shard(qv), mk_shard(kv_), mk_shard(vv), layout)
        ref = split_varlen_input(full, seqs, world, rank, "HND")              # zero padded like the shards: padded rows must be zeros
        errs[f"varlen_ring/{layout}"] = float((_from_layout(out, layout).float() - ref).abs().max())
    return errs


def _cases_world4(rank, world):
    from flashinfer_b200.parallel_attention import ParallelAttention, UnevenCPConfig, uneven_cp_config

    errs = {}
    g = torch.Generator().manual_seed(1)
    H, S, D = 4, 64, 16
    q, k, v = (torch.randn(H, S, D, generator=g) for _ in range(3))
    shard = lambda t: t.chunk(world, dim=1)[rank]  # noqa: E731
    ring_g, uly_g = _groups(world, 2, 2)
    assert dist.get_process_group_ranks(uly_g) == [rank // 2 * 2, rank // 2 * 2 + 1] and dist.get_process_group_ranks(ring_g) == [rank % 2, rank % 2 + 2]
    for layout in ("HND", "NHD"):
        out = ParallelAttention("sm100", uly_g, ring_g, fuse_qkv=True).run(*(_to_layout(shard(t), layout) for t in (q, k, v)), layout)
        errs[f"2d/{layout}"] = float((_from_layout(out, layout).float() - shard(_sdpa(q, k, v))).abs().max())
    pad = world - 1
    real = S - pad
    mine = S // world - (pad if rank == world - 1 else 0)
    lens = uneven_cp_config(real, S, mine, uly_g, ring_g)
    assert lens.tolist() == [2 * (S // world), 2 * (S // world) - pad]
    cfg = UnevenCPConfig(real, S, lens)
    out = ParallelAttention("sm100", uly_g, ring_g, uneven_cp_config=cfg).run(shard(q), shard(k), shard(v), "HND")
    ref = torch.zeros(H, S, D)
    ref[:, :real] = _sdpa(q[:, :real], k[:, :real], v[:, :real])
    errs["2d/uneven"] = float((out.float() - shard(ref)).abs().max())
    with pytest.raises(NotImplementedError):
        ParallelAttention("sm100", uly_g, ring_g).run(shard(q), shard(k), shard(v), "HND", is_causal=True)
    return errs


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        results[rank] = (_cases_world2 if world == 2 else _cases_world4)(rank, world)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_parallel_attention_gloo(world):
    results = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    assert len(results) == world
    worst = {name: max(results[r][name] for r in range(world)) for name in results[0]}
    bad = {n: e for n, e in worst.items() if not e < 2e-4}
    assert not bad, bad
    assert len(worst) == (21 if world == 2 else 3)


def test_helpers_single_process():
    from flashinfer_b200.parallel_attention import AttentionOpManager, split_varlen_input
    from flashinfer_b200.parallel_attention.parallel_wrapper import (get_kv_rank, ring_fwd_out_correction, ring_fwd_softmax_lse_correction)
    from flashinfer_b200.parallel_attention.utils import convert_output_layout, convert_qkv_layout

    x = torch.arange(2 * 10 * 1, dtype=torch.float32).view(2, 10, 1)           # HND, sequences of 7 and 3 tokens
    shards = [split_varlen_input(x, [7, 3], 3, r, "HND") for r in range(3)]
    assert [s.shape[1] for s in shards] == [4, 4, 4]                             # ceil(7 / 3) + ceil(3 / 3)
    assert shards[0][0, :, 0].tolist() == [0, 1, 2, 7] and shards[1][0, :, 0].tolist() == [3, 4, 5, 8] and shards[2][0, :, 0].tolist() == [6, 9, 0, 0]
    assert torch.equal(split_varlen_input(x.transpose(0, 1), [7, 3], 3, 1, "NHD"), shards[1].transpose(0, 1))
    q = torch.randn(3, 5, 4)
    qn, kn, vn = convert_qkv_layout(q, q, q, "HND", "NHD")
    assert qn.shape == (5, 3, 4) and torch.equal(convert_output_layout(qn, "NHD", "HND"), q)
    with pytest.raises(ValueError):
        convert_qkv_layout(q, q, q, "HND", "BHSD")
    assert [get_kv_rank(4, 1, i) for i in range(4)] == [1, 0, 3, 2]
    # merging two partial softmax results reproduces the joint one
    lg = torch.randn(6, 9)
    vals = torch.randn(9, 5)
    full = torch.softmax(lg, -1) @ vals
    a, b = slice(0, 4), slice(4, 9)
    out = (torch.softmax(lg[:, a], -1) @ vals[a]).clone()
    lse = torch.logsumexp(lg[:, a], -1)
    ring_fwd_out_correction(out, torch.softmax(lg[:, b], -1) @ vals[b], lse, torch.logsumexp(lg[:, b], -1))
    ring_fwd_softmax_lse_correction(lse, torch.logsumexp(lg[:, b], -1))
    torch.testing.assert_close(out, full, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(lse, torch.logsumexp(lg, -1), atol=1e-5, rtol=1e-5)
    empty = torch.full((6,), float("-inf"))
    ring_fwd_out_correction(out, torch.zeros_like(out), lse, empty)              # an empty block changes nothing
    torch.testing.assert_close(out, full, atol=1e-5, rtol=1e-5)

    @AttentionOpManager.register_attn("unit-test-backend")
    class Mine:
        def __call__(self, **kw):
            return "called"

    assert "unit-test-backend" in AttentionOpManager.get_registered_types() and AttentionOpManager.get_impl("unit-test-backend")() == "called"
    with pytest.raises(ValueError):
        AttentionOpManager.get_impl("nope")
    AttentionOpManager._attn_registry.pop("unit-test-backend")
