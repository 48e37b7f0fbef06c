"""Tensor / expert parallel runs of the configurable engine on 2 gloo ranks equal the single-process run (heads, MLP columns and whole
experts are sharded; weights are slices of the same full-model draw)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flashinfer_b200.models import TransformerConfig, TransformerDecodeEngine
from flashinfer_b200.models.serving import PagedKVAllocator, generate

PROMPTS = [[3, 14, 15, 92, 65, 35], [79, 32, 38]]


def _run(cfg, group):
    eng = TransformerDecodeEngine(cfg, max_batch=2, max_pages=16, page_size=4, device="cpu", dtype=torch.bfloat16, seed=6, tp_group=group)
    alloc = PagedKVAllocator(16, 4, seed=2)
    for r, p in enumerate(PROMPTS):
        alloc.add_request(r)
        alloc.grow(r, len(p))
    qo = torch.tensor([0, len(PROMPTS[0]), len(PROMPTS[0]) + len(PROMPTS[1])], dtype=torch.int32)
    eng.prefill(torch.tensor(PROMPTS[0] + PROMPTS[1]), qo, *alloc.tables([0, 1]))
    prefill_logits = eng.logits.float().clone()                              # same inputs on every configuration
    for r in (0, 1):
        alloc.grow(r, 1)
    eng.plan(*alloc.tables([0, 1]))
    eng.tokens.copy_(torch.tensor([7, 11]))                                   # a fixed continuation: decode logits are comparable too
    eng.step()
    eng2 = TransformerDecodeEngine(cfg, max_batch=2, max_pages=16, page_size=4, device="cpu", dtype=torch.bfloat16, seed=6, tp_group=group)
    toks = generate(eng2, PROMPTS, 3, PagedKVAllocator(16, 4, seed=2))
    return toks, torch.cat([prefill_logits, eng.logits.float()])


def _worker(rank, world, port, family, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = (TransformerConfig() if family == "plain" else getattr(TransformerConfig, family)()).tiny()
        toks, logits = _run(cfg, dist.new_group(list(range(world))))
        out[rank] = (toks, logits)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("family", ["plain", "qwen3_30b_a3b", "gemma2_9b"])
def test_tp2_matches_single_process(family):
    cfg = (TransformerConfig() if family == "plain" else getattr(TransformerConfig, family)()).tiny()
    ref_toks, ref_logits = _run(cfg, None)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, family, out), nprocs=2, join=True)
    for rank in range(2):
        toks, logits = out[rank]
        cos = torch.nn.functional.cosine_similarity(logits.flatten(), ref_logits.flatten(), dim=0)
        assert cos > 0.999, (family, rank, float(cos))
    assert out[0][0] == out[1][0]                                            # the ranks agree on every generated token
    for r, toks in enumerate(out[0][0]):                                     # the first greedy token is the single-process argmax, near-ties aside
        row = ref_logits[r]
        assert float(row[toks[0]]) >= float(row.max()) - 0.02 * float(row.max() - row.min())


def test_sharding_requirements():
    class _G:                                                                 # a stand-in group is enough to hit the divisibility check
        pass

    cfg = TransformerConfig(num_qo_heads=6, num_kv_heads=3).tiny()
    cfg.num_qo_heads, cfg.num_kv_heads = 6, 3
    orig_ws, orig_rk = dist.get_world_size, dist.get_rank
    dist.get_world_size, dist.get_rank = (lambda g=None: 2), (lambda g=None: 0)
    try:
        with pytest.raises(ValueError):
            TransformerDecodeEngine(cfg, 1, 4, 4, "cpu", torch.bfloat16, tp_group=_G())
    finally:
        dist.get_world_size, dist.get_rank = orig_ws, orig_rk


def test_parallel_helpers():
    from flashinfer_b200 import parallel

    w = torch.arange(24.0).view(6, 4)                                         # two stacked blocks of three rows
    assert parallel.shard_rows(w, 1, 3, blocks=2)[:, 0].tolist() == [4.0, 16.0] and parallel.shard_rows(w, 0, 2)[:, 0].tolist() == [0.0, 4.0, 8.0]
    assert torch.equal(torch.cat([parallel.shard_cols(w, r, 2) for r in range(2)], -1), w)
    with pytest.raises(ValueError):
        parallel.shard_rows(w, 0, 4)
    with pytest.raises(ValueError):
        parallel.shard_cols(w, 0, 3)
    x = torch.randn(3, 5)
    assert parallel.all_reduce_fp32(x, None) is x
