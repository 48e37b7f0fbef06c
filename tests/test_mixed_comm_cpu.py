"""comm.mixed_comm in the reference's conventions (TP x DP topology, ``run_mixed_comm(op, handler, x_in, x_out, mode)``) on gloo ranks:
every op against its definition evaluated from the known per-rank inputs."""
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


def _x(rank, rows=4, cols=3):
    return torch.arange(rows * cols, dtype=torch.float32).view(rows, cols) * (rank + 1) + rank


def _check(rank, world, tp, dp):
    from flashinfer_b200.comm.mixed_comm import MixedCommHandler, MixedCommMode, MixedCommOp, run_mixed_comm

    h = MixedCommHandler(rank, world, rank, world, 0, 1, tp, dp, None, None, torch.float32, torch.device("cpu"), use_autotune=rank % 2 == 0)
    p = h.para_info
    assert (p.tp_rank, p.dp_rank) == (rank % p.tp_size, rank // p.tp_size) and p.tp_size * p.dp_size == world
    tp_peers = [r for r in range(world) if r // p.tp_size == p.dp_rank]
    dp_peers = [r for r in range(world) if r % p.tp_size == p.tp_rank]
    assert p.tp_groups()[p.dp_rank] == tp_peers and p.dp_groups()[p.tp_rank] == dp_peers
    errs = {}
    want = {
        MixedCommOp.ALLREDUCE: lambda: sum(_x(r) for r in tp_peers),
        MixedCommOp.ALLGATHER: lambda: torch.cat([_x(r) for r in dp_peers]),
        MixedCommOp.REDUCESCATTER: lambda: sum(_x(r, 4 * p.dp_size) for r in dp_peers).unflatten(0, (p.dp_size, -1))[p.dp_rank],
        MixedCommOp.ALLREDUCE_ALLGATHER: lambda: torch.cat([sum(_x(r) for r in range(world) if r // p.tp_size == d) for d in range(p.dp_size)]),
        MixedCommOp.REDUCESCATTER_ALLREDUCE: lambda: sum(_x(r, 4 * p.dp_size) for r in range(world)).unflatten(0, (p.dp_size, -1))[p.dp_rank],
    }
    for op in MixedCommOp:
        rows = 4 * p.dp_size if op in (MixedCommOp.REDUCESCATTER, MixedCommOp.REDUCESCATTER_ALLREDUCE) else 4
        if op not in h.valid_op_list:
            try:
                run_mixed_comm(op, h, _x(rank, rows))
                errs[f"{op.name}:accepted"] = 1.0
            except ValueError:
                pass
            continue
        got = run_mixed_comm(op, h, _x(rank, rows))
        errs[op.name] = float((got - want[op]()).abs().max())
        buf = torch.empty_like(got)
        assert h.run(op, _x(rank, rows), buf, MixedCommMode.NCCL_ONE) is buf
        errs[op.name + ":out"] = float((buf - want[op]()).abs().max())
    for bad in (lambda: run_mixed_comm(h.valid_op_list[0], h, torch.zeros(4)), lambda: run_mixed_comm(h.valid_op_list[0], h, torch.zeros(4, 3, dtype=torch.float16)),
                lambda: run_mixed_comm(h.valid_op_list[0], h, _x(rank), None, MixedCommMode.FUSED_OPT_WAITS_MC)):
        try:
            bad()
            errs["bad-input-accepted"] = 1.0
        except ValueError:
            pass
    h.shutdown()
    return errs


def _worker(rank, world, port, tp, dp, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        results[rank] = _check(rank, world, tp, dp)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,tp,dp,n_ops", [(4, 2, 2, 5), (2, 2, None, 1), (2, None, 2, 2), (4, None, None, 1)])
def test_mixed_comm_gloo(world, tp, dp, n_ops):
    results = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), tp, dp, results), nprocs=world, join=True)
    assert len(results) == world
    for r in range(world):
        assert len(results[r]) == 2 * n_ops, results[r]
        assert all(e == 0.0 for e in results[r].values()), (r, results[r])


def test_parallel_info_two_nodes():
    from flashinfer_b200.comm.mixed_comm import ParallelInfo

    p = ParallelInfo(13, 16, 5, 8, 1, 2, 4, 2, 1, 2)             # 2 nodes x 8 GPUs, TP4 x DP2 inside a node, DP across nodes
    assert (p.tp_size, p.dp_size, p.tp_rank, p.dp_rank) == (4, 4, 1, 3)
    assert p.tp_groups()[3] == [12, 13, 14, 15] and p.dp_groups()[1] == [1, 5, 9, 13]
    assert sorted(sum(p.tp_groups(), [])) == list(range(16)) and sorted(sum(p.dp_groups(), [])) == list(range(16))
    q = ParallelInfo(13, 16, 5, 8, 1, 2, None, None, None, None)  # defaults: everything is TP
    assert (q.tp_size, q.dp_size, q.tp_rank) == (16, 1, 13) and q.use_tp and not q.use_dp and not q.use_mixed
