"""All-gather / reduce-scatter / all-to-all (DCP) kernels vs torch.distributed oracles on spawned ranks."""
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, backend, errs):
    from flashinfer_b200.comm import (NVLSCollectives, decode_cp_a2a_allocate_mnnvl_workspace, decode_cp_a2a_alltoall,
                                      decode_cp_a2a_init_workspace)

    cuda = backend == "nccl"
    if cuda:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                device_id=torch.device("cuda", rank))
    else:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    dev = torch.device("cuda", rank) if cuda else torch.device("cpu")
    dtype = torch.bfloat16 if cuda else torch.float32
    try:
        worst = 0.0
        for use_nvls in ([True, False] if cuda else [True]):
            coll = NVLSCollectives(None, 8 << 20, use_nvls=use_nvls)
            for it, n in enumerate([8, 300, 1024]):
                torch.manual_seed(10 * it + rank)
                x = torch.randn(n, 512, dtype=dtype, device=dev)
                ref = [torch.empty_like(x) for _ in range(world)]
                dist.all_gather(ref, x)
                got = coll.all_gather(x)
                worst = max(worst, float((got.float() - torch.cat(ref).float()).abs().max()))
                y = torch.randn(world * n, 512, dtype=dtype, device=dev)
                full = y.float().clone()
                dist.all_reduce(full)
                got = coll.reduce_scatter(y)
                worst = max(worst, float((got.float() - full[rank * n:(rank + 1) * n]).abs().max()) / 8.0)
                z = torch.randn(n, world, 128, dtype=dtype, device=dev)
                got = coll.all_to_all(z)
                zs = [torch.empty_like(z) for _ in range(world)]
                dist.all_gather(zs, z)
                exp = torch.stack([zs[i][:, rank] for i in range(world)], 1)
                worst = max(worst, float((got.float() - exp.float()).abs().max()))
        if cuda:
            ws = decode_cp_a2a_allocate_mnnvl_workspace(None)
            decode_cp_a2a_init_workspace(ws, rank, world)
            po = torch.randn(4, 16, world, 128, dtype=dtype, device=dev)
            st = torch.randn(4, 16, world, 2, dtype=torch.float32, device=dev)
            o, s = decode_cp_a2a_alltoall(po, st, ws, rank, world)
            pos = [torch.empty_like(po) for _ in range(world)]
            sts = [torch.empty_like(st) for _ in range(world)]
            dist.all_gather(pos, po)
            dist.all_gather(sts, st)
            worst = max(worst, float((o - torch.stack([pos[i][..., rank, :] for i in range(world)], -2)).abs().max()))
            worst = max(worst, float((s - torch.stack([sts[i][..., rank, :] for i in range(world)], -2)).abs().max()))
            torch.cuda.synchronize()
        errs[rank] = worst
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_collectives_gloo():
    errs = mp.Manager().dict()
    mp.spawn(_worker, args=(2, _free_port(), "gloo", errs), nprocs=2, join=True)
    assert max(errs.values()) < 1e-5, dict(errs)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 8])
def test_collectives_nvlink(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    errs = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), "nccl", errs), nprocs=world, join=True)
    assert max(errs.values()) < 2e-2, dict(errs)
