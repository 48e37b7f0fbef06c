"""Fused GEMM+all-reduce and all-gather+GEMM kernels vs NCCL + torch oracles (reference
tests/gemm/test_cute_dsl_gemm_allreduce_two_shot.py, tests/comm/test_all_gather_matmul.py)."""
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


def _worker(rank, world, port, errs):
    from flashinfer_b200.comm import AllGatherMatmul, GemmAllReduce

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        worst = 0.0
        N, K = 4096, 2048
        for use_nvls in (True, False):
            gar = GemmAllReduce(None, 2048, N, torch.bfloat16, use_nvls=use_nvls)
            for it, (M, two_shot) in enumerate([(1, False), (64, False), (200, False), (1024, True), (64, True), (2048, None)]):
                torch.manual_seed(17 * it + rank)
                a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
                w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
                ref = (a.float() @ w.float().t()).bfloat16().float()
                dist.all_reduce(ref)
                got = gar(a, w, two_shot=two_shot)
                torch.cuda.synchronize()
                worst = max(worst, float((got.float() - ref).abs().max() / ref.abs().max()))
            # GEMM -> reduce-scatter (+ residual + RMSNorm) in the same kernel (BASELINE config 5)
            for it, M in enumerate([world * 8, world * 128, 2048, world * 72]):
                torch.manual_seed(91 * it + rank)
                a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
                w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
                rpr = M // world
                res = torch.randn(rpr, N, device="cuda").bfloat16()
                gamma = (1 + 0.1 * torch.randn(N, device="cuda")).bfloat16()
                full = (a.float() @ w.float().t()).bfloat16().float()
                dist.all_reduce(full)
                mine = full[rank * rpr:(rank + 1) * rpr]
                got = gar.reduce_scatter(a, w)
                torch.cuda.synchronize()
                worst = max(worst, float((got.float() - mine).abs().max() / full.abs().max()))
                if M == 2048:  # every schedule of the prefill-size path: one kernel, full GEMM + full pull, 2 / 4 pipeline chunks
                    for sched in (0, 1, 2, 4):
                        normed, shard = gar.reduce_scatter(a, w, residual=res, rms_weight=gamma, eps=1e-5, pipelined=sched)
                        torch.cuda.synchronize()
                        x = mine + res.float()
                        worst = max(worst, float((shard.float() - x).abs().max() / x.abs().max()))
                for _ in range(2):  # twice: the sum-of-squares scratch must be left clean
                    normed, shard = gar.reduce_scatter(a, w, residual=res, rms_weight=gamma, eps=1e-5)
                    torch.cuda.synchronize()
                    x = mine + res.float()
                    ref_n = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5) * gamma.float()
                    worst = max(worst, float((shard.float() - x).abs().max() / x.abs().max()))
                    worst = max(worst, float((normed.float() - ref_n).abs().max() / ref_n.abs().max()))
            agm = AllGatherMatmul(None, 512, K, torch.bfloat16, use_nvls=use_nvls)
            for it, Ml in enumerate([128, 512, 256]):
                torch.manual_seed(31 * it + rank)
                x = (torch.randn(Ml, K, device="cuda") * 0.5).bfloat16()
                torch.manual_seed(5 + it)
                w = (torch.randn(1024, K, device="cuda") / K ** 0.5).bfloat16()
                xs = [torch.empty_like(x) for _ in range(world)]
                dist.all_gather(xs, x)
                ref = torch.cat(xs).float() @ w.float().t()
                got, gathered = agm(x, w, return_gathered=True)
                torch.cuda.synchronize()
                worst = max(worst, float((got.float() - ref).abs().max() / ref.abs().max()))
                worst = max(worst, float((gathered.float() - torch.cat(xs).float()).abs().max()))
        errs[rank] = worst
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_gemm_comm_fused(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    errs = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), errs), nprocs=world, join=True)
    assert max(errs.values()) < 2e-2, dict(errs)
