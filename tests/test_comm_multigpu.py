"""Multi-GPU tests of the in-kernel NVLink all-reduce fusion (spawned ranks, like the reference's
tests/comm/test_trtllm_allreduce_fusion.py:27-110)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, use_nvls, results):
    import flashinfer_b200 as fi  # noqa: F401
    from flashinfer_b200.comm import TPCommunicator
    from flashinfer_b200 import reference

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        hidden = 4096
        comm = TPCommunicator(None, max_tokens=2048, hidden=hidden, dtype=torch.bfloat16, use_nvls=use_nvls)
        for tokens, two_shot in [(1, False), (64, False), (128, False), (300, True), (2048, True)]:
            for it in range(3):
                torch.manual_seed(100 * it + rank)
                part = torch.randn(tokens, hidden, device="cuda", dtype=torch.bfloat16)
                torch.manual_seed(7 + it)
                res0 = torch.randn(tokens, hidden, device="cuda", dtype=torch.bfloat16)
                w = torch.randn(hidden, device="cuda", dtype=torch.bfloat16)
                # oracle: NCCL all-reduce in fp32 + torch add + rmsnorm
                full = part.float().clone()
                dist.all_reduce(full)
                r_ref = (full + res0.float()).to(torch.bfloat16)
                y_ref = reference.rmsnorm_ref(r_ref, w, 1e-5)
                buf = comm.gemm_out(tokens)
                buf.copy_(part)
                res = res0.clone()
                out = comm.allreduce_add_rmsnorm(buf, res, w, 1e-5, two_shot=two_shot)
                torch.cuda.synchronize()
                torch.testing.assert_close(out.float(), y_ref.float(), rtol=3e-2, atol=3e-2)
                if two_shot:
                    torch.testing.assert_close(res[rank::world].float(), r_ref[rank::world].float(), rtol=2e-2, atol=2e-2)
                else:
                    torch.testing.assert_close(res.float(), r_ref.float(), rtol=2e-2, atol=2e-2)
        # plain all-reduce + small fp32 all-reduce / argmax gather
        x = torch.randn(64, hidden, device="cuda", dtype=torch.bfloat16)
        ref = x.float().clone()
        dist.all_reduce(ref)
        got = comm.all_reduce(x)
        torch.testing.assert_close(got.float(), ref, rtol=3e-2, atol=3e-2)
        val = torch.randn(16, device="cuda")
        idx = torch.arange(16, device="cuda") + 1000 * rank
        allv = [torch.empty_like(val) for _ in range(world)]
        dist.all_gather(allv, val)
        best = torch.stack(allv).argmax(0)
        got = comm.argmax_gather(val, idx)
        assert got.tolist() == (torch.arange(16, device="cuda") + 1000 * best).tolist()
        # CUDA-graph replay safety of the epoch barriers
        buf = comm.gemm_out(64)
        res = torch.zeros(64, hidden, device="cuda", dtype=torch.bfloat16)
        w = torch.ones(hidden, device="cuda", dtype=torch.bfloat16)
        out = torch.empty(64, hidden, device="cuda", dtype=torch.bfloat16)
        buf.fill_(float(rank + 1))
        comm.allreduce_add_rmsnorm(buf, res, w, 1e-5, out=out)
        torch.cuda.synchronize()
        dist.barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            comm.allreduce_add_rmsnorm(buf, res, w, 1e-5, out=out)
        for _ in range(5):
            g.replay()
        torch.cuda.synchronize()
        assert torch.isfinite(out.float()).all()
        results[rank] = "ok" + (" nvls" if comm.use_nvls else " p2p")
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("use_nvls", [True, False])
def test_allreduce_fusion_multigpu(use_nvls):
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if world < 4 else (4 if world < 8 else 8)
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, use_nvls, results), nprocs=world, join=True)
    assert all(v.startswith("ok") for v in results.values()) and len(results) == world, dict(results)
