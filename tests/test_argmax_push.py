"""Vocabulary-parallel greedy sampling in one kernel (csrc/comm/allreduce.cu: argmax_push_kernel) against
all_gather + torch.argmax; eager calls and CUDA-graph replays (the epoch / parity rotation must survive both)."""
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, errs):
    import torch.distributed as dist

    from flashinfer_b200.comm import TPCommunicator

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        comm = TPCommunicator(dist.group.WORLD, max_tokens=64, hidden=1024, dtype=torch.bfloat16)
        bad = 0
        for it, (rows, shard, dt) in enumerate([(64, 16032, torch.bfloat16), (1, 1000, torch.float32), (33, 4099, torch.float16),
                                                (64, 16032, torch.bfloat16), (7, 515, torch.bfloat16)]):
            torch.manual_seed(31 * it + rank)
            logits = torch.randn(rows, shard, device="cuda").to(dt)
            if it == 3:  # ties across ranks and inside a shard: the lowest global index must win everywhere
                logits.fill_(0.0)
                logits[:, 5] = 3.0
                logits[:, 9] = 3.0
            out = comm.argmax_logits(logits, rank * shard)
            val = torch.empty(rows, device="cuda")
            comm.argmax_logits(logits, rank * shard, out=out, out_val=val)
            allv = [torch.empty_like(logits) for _ in range(world)]
            dist.all_gather(allv, logits)
            full = torch.cat(allv, 1).float()
            ref = full.argmax(-1)
            # torch.argmax does not promise the first index on ties: compare values, and indices where the maximum is unique
            mx = full.max(-1).values
            uniq = (full == mx[:, None]).sum(-1) == 1
            bad += int((val != mx).sum()) + int((out[uniq] != ref[uniq]).sum())
            if it == 3:
                bad += int((out != 5).sum())
        # CUDA graph: 3 calls per replay, 4 replays
        logits = torch.randn(64, 2048, device="cuda").bfloat16()
        out = torch.empty(64, dtype=torch.int64, device="cuda")
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            comm.argmax_logits(logits, rank * 2048, out=out)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(3):
                comm.argmax_logits(logits, rank * 2048, out=out)
        for _ in range(4):
            g.replay()
        torch.cuda.synchronize()
        allv = [torch.empty_like(logits) for _ in range(world)]
        dist.all_gather(allv, logits)
        full = torch.cat(allv, 1).float()
        bad += int((full.gather(1, out[:, None])[:, 0] != full.max(-1).values).sum())
        errs[rank] = bad
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_argmax_push(world):
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    errs = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), errs), nprocs=world, join=True)
    assert len(errs) == world and max(errs.values()) == 0, dict(errs)


@pytest.mark.parametrize("rows,v,dt", [(64, 128256, torch.bfloat16), (3, 1000, torch.float32), (64, 16032, torch.float16)])
def test_local_argmax(rows, v, dt):
    """Single-GPU flavour (group of one) used by the decode engine for greedy sampling."""
    from flashinfer_b200.comm.allreduce import local_argmax

    torch.manual_seed(rows)
    x = torch.randn(rows, v, device="cuda").to(dt)
    x[0, 5] = x[0, 9] = 100.0  # tie: lowest index wins
    val = torch.empty(rows, device="cuda")
    for _ in range(3):  # epoch / parity rotation
        out = local_argmax(x, out_val=val)
    assert int(out[0]) == 5
    assert torch.equal(val, x.float().max(-1).values)
    assert torch.equal(x.float().gather(1, out[:, None])[:, 0], val)
