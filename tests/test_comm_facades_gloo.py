"""nvshmem-style all-to-all facade over gloo (2 processes, CPU): host-side logic of comm.nvshmem."""
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ok):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        from flashinfer_b200.comm import nvshmem
        from flashinfer_b200.comm.mnnvl import TorchDistBackend

        assert nvshmem.my_pe() == rank and nvshmem.n_pes() == world
        src = torch.arange(world * 3, dtype=torch.float32) + 100 * rank
        dst = torch.empty_like(src)
        nvshmem.alltoall(dst, src)
        exp = torch.cat([torch.arange(rank * 3, rank * 3 + 3, dtype=torch.float32) + 100 * r for r in range(world)])
        assert torch.equal(dst, exp), (dst, exp)
        be = TorchDistBackend()
        assert be.Get_rank() == rank and be.allgather(rank) == list(range(world)) and be.bcast("x" if rank == 0 else None) == "x"
        be.barrier()
        ok[rank] = True
    finally:
        dist.destroy_process_group()


def test_nvshmem_facade_gloo():
    ok = mp.Manager().dict()
    mp.spawn(_worker, args=(2, _free_port(), ok), nprocs=2, join=True)
    assert all(ok.get(r) for r in range(2))
