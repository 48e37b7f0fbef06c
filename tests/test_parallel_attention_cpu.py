"""Ulysses / Ring attention on 2 gloo ranks vs single-process full attention (reference tests/attention/
test_parallel_attention*.py strategy: every rank checks its shard of the global result)."""
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


def _worker(rank, world, port, mode, causal, q, k, v, ref, errs):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flashinfer_b200.parallel_attention import ParallelAttention

        S = q.shape[0] // world
        sl = slice(rank * S, (rank + 1) * S)
        g = dist.new_group(list(range(world)))
        pa = ParallelAttention(ulysses_group=g if mode == "ulysses" else None, ring_group=g if mode == "ring" else None,
                               fuse_qkv=(mode == "ulysses"))
        out = pa.run(q[sl].clone(), k[sl].clone(), v[sl].clone(), "NHD", is_causal=causal)
        errs[rank] = float((out.float() - ref[sl]).abs().max())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode,causal", [("ulysses", False), ("ulysses", True), ("ring", False), ("ring", True)])
def test_parallel_attention_gloo(mode, causal):
    from flashinfer_b200.prefill import single_prefill_with_kv_cache

    torch.manual_seed(0)
    S, H, D = 64, 4, 32
    q, k, v = (torch.randn(S, H, D) for _ in range(3))
    ref = single_prefill_with_kv_cache(q, k, v, causal=causal).float()
    errs = mp.Manager().dict()
    mp.spawn(_worker, args=(2, _free_port(), mode, causal, q, k, v, ref, errs), nprocs=2, join=True)
    assert max(errs.values()) < 1e-4, dict(errs)
