"""Legacy prepare + comm MoE all-to-all (flashinfer_b200.comm.trtllm_alltoall): index preparation and local gather against
brute-force oracles (CPU), the native indexed all-to-all-v on one GPU, and the full dispatch -> "experts" -> combine round trip
over 2 / 4 / 8 GPUs.  Port of reference tests/comm/test_trtllm_alltoall.py (:109 single GPU, :323 prepare_indices, :431 local
gather, :536 prepare) and tests/comm/test_mnnvl_moe_alltoall.py."""
import socket

import pytest
import torch

from flashinfer_b200.comm import trtllm_alltoall as ta


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("ep_rank,ep_size,top_k,max_tok,real", [(0, 2, 3, 17, False), (3, 8, 8, 40, True), (1, 4, 2, 5, True), (2, 4, 6, 33, False)])
def test_prepare_indices_oracle(ep_rank, ep_size, top_k, max_tok, real):
    torch.manual_seed(ep_rank * 7 + ep_size)
    counts = [int(torch.randint(1, max_tok + 1, (1,))) for _ in range(ep_size)] if real else [max_tok] * ep_size
    cum = torch.tensor(counts).cumsum(0).int()
    total = int(cum[-1])
    g = torch.randint(0, ep_size, (total, top_k), dtype=torch.int32)
    out = ta.moe_comm_prepare_indices(g, cum if real else None, max_tok, 64, top_k, ep_rank, ep_size)
    gather, send_cum, send_idx, recv_cum, recv_idx, backward = out
    lo = int(cum[ep_rank - 1]) if ep_rank else 0
    hi = int(cum[ep_rank])
    # oracle: python loops
    send, back = [[] for _ in range(ep_size)], [[] for _ in range(ep_size)]
    for t in range(lo, hi):
        for r in range(ep_size):
            ks = [k for k in range(top_k) if int(g[t, k]) == r]
            if ks:
                send[r].append(t - lo)
                back[r].append((t - lo) * top_k + ks[0])
    recv = [[] for _ in range(ep_size)]
    for s in range(ep_size):
        slo = int(cum[s - 1]) if s else 0
        for t in range(slo, int(cum[s])):
            if (g[t] == ep_rank).any():
                recv[s].append(t)
    assert send_idx.tolist() == [t for r in send for t in r]
    assert backward.tolist() == [t for r in back for t in r]
    assert send_cum.tolist() == torch.tensor([len(r) for r in send]).cumsum(0).tolist()
    assert recv_cum.tolist() == torch.tensor([len(r) for r in recv]).cumsum(0).tolist()
    assert gather.tolist() == [t for r in recv for t in r]
    assert recv_idx.tolist() == list(range(sum(len(r) for r in recv)))


def test_local_gather_and_cpu_comm():
    torch.manual_seed(0)
    ep_size, top_k, max_tok, E = 4, 3, 9, 32
    cum = torch.randint(0, max_tok + 1, (ep_size,)).cumsum(0).int()
    n = int(cum[-1])
    alloc = max_tok * ep_size
    idx = torch.randint(0, alloc, (alloc,), dtype=torch.int32)
    ids = torch.randint(0, E, (alloc, top_k), dtype=torch.int32)
    sc = torch.rand(alloc, top_k)
    lid, lsc = torch.empty(alloc, top_k, dtype=torch.int32), torch.empty(alloc, top_k)
    ta.moe_local_gather(cum, idx, ids, sc, lid, lsc, max_tok, E, top_k, 1, ep_size)
    assert torch.equal(lid[:n], ids[idx[:n].long()]) and (lid[n:] == E).all()
    assert torch.equal(lsc[:n], sc[idx[:n].long()]) and (lsc[n:] == 0).all()
    # single-rank moe_comm on CPU tensors = indexed copy
    x = torch.randn(20, 16)
    out = torch.zeros(30, 16)
    si, ri = torch.randperm(20)[:11].int(), torch.randperm(30)[:11].int()
    ta.moe_comm(x, torch.tensor([11], dtype=torch.int32), si, out, torch.tensor([11], dtype=torch.int32), ri, torch.zeros(8), 0, 1)
    ref = torch.zeros(30, 16)
    ref[ri.long()] = x[si.long()]
    assert torch.equal(out, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("n_in,n_out,dim,cnt,dtype", [(1024, 2048, 512, 700, torch.bfloat16), (16, 16, 8, 16, torch.float32), (4096, 4096, 7168, 3000, torch.bfloat16)])
def test_moe_comm_single_gpu(n_in, n_out, dim, cnt, dtype):
    x = torch.randn(n_in, dim, device="cuda").to(dtype)
    out = torch.zeros(n_out, dim, device="cuda", dtype=dtype)
    si = torch.randperm(n_in, device="cuda")[:cnt].int()
    ri = torch.randperm(n_out, device="cuda")[:cnt].int()
    cum = torch.tensor([cnt], dtype=torch.int32, device="cuda")
    ws = torch.zeros(1, ta.get_moe_commworkspace_size_per_rank(1) // 8, dtype=torch.uint64, device="cuda")
    for _ in range(2):  # twice: barrier epochs / staging reuse
        out.zero_()
        ta.moe_comm(x, cum, si, out, cum, ri, ws, 0, 1)
    ref = torch.zeros_like(out)
    ref[ri.long()] = x[si.long()]
    assert torch.equal(out, ref)


def _worker(rank, world, port, errs):
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        from flashinfer_b200.comm import Mapping

        E, K, H, max_tok = 8 * world, 4, 1024, 96
        mapping = Mapping(world, rank, tp_size=world, moe_ep_size=world, moe_tp_size=1)
        ws = ta.MnnvlMoe.get_moe_workspaces(mapping)
        worst = 0.0
        for it in range(3):
            torch.manual_seed(50 * it + rank)
            T = [max_tok, 1 + 13 * rank % max_tok, 37][it]
            x = torch.randn(T, H, device="cuda").bfloat16()
            ids = torch.stack([torch.randperm(E, device="cuda")[:K] for _ in range(T)]).int()
            sc = torch.rand(T, K, device="cuda")
            info, lids, lsc, _ = ta.MnnvlMoe.mnnvl_moe_alltoallv_prepare_without_allgather(ids, sc, None, None, max_tok, rank, world, E, E, K)
            recv = ta.MnnvlMoe.mnnvl_moe_alltoallv(x, info, ws, rank, world)
            n_recv = int(info.recv_rank_count_cumsum[-1])
            # "experts": every local expert e scales the token by (e + 1); weighted by the routing scale of the slots that live here
            epr = E // world
            here = (lids[:n_recv] >= rank * epr) & (lids[:n_recv] < (rank + 1) * epr)
            coef = (torch.where(here, (lids[:n_recv] + 1).float() * lsc[:n_recv], torch.zeros_like(lsc[:n_recv]))).sum(-1)
            y = torch.zeros_like(recv)
            y[:n_recv] = (recv[:n_recv].float() * coef[:, None]).bfloat16()
            out = ta.MnnvlMoe.mnnvl_moe_alltoallv_combine(y, info, ws, rank, world, K, T)
            # oracle: per-rank partial sums are rounded to bf16 before the final sum over ranks
            ref = torch.zeros(T, H, device="cuda")
            for r in range(world):
                m = (ids >= r * epr) & (ids < (r + 1) * epr)
                c = torch.where(m, (ids + 1).float() * sc, torch.zeros_like(sc)).sum(-1)
                ref += (x.float() * c[:, None]).bfloat16().float() * m.any(-1)[:, None]
            worst = max(worst, float((out.float() - ref).abs().max() / ref.abs().max()))
        errs[rank] = worst
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_legacy_alltoall_round_trip(world):
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    errs = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), errs), nprocs=world, join=True)
    assert len(errs) == world and max(errs.values()) < 2e-2, dict(errs)


def _gloo_worker(rank, world, port, errs):
    """Host-side logic of the legacy protocol across ranks, on CPU tensors over gloo: prepare (all-gathered routing tables, index
    lists), dispatch, "experts", combine == a local oracle."""
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        E, K, H, max_tok = 4 * world, 3, 16, 12
        epr = E // world
        worst = 0.0
        for it in range(3):
            torch.manual_seed(10 * it + rank)
            T = [max_tok, 1 + 5 * rank, 7][it]
            x = torch.randn(T, H)
            ids = torch.stack([torch.randperm(E)[:K] for _ in range(T)]).int()
            sc = torch.rand(T, K)
            info, lids, lsc, _ = ta.MnnvlMoe.mnnvl_moe_alltoallv_prepare_without_allgather(ids, sc, None, None, max_tok, rank, world, E, E, K)
            ws = torch.zeros(8)
            recv = ta.MnnvlMoe.mnnvl_moe_alltoallv(x, info, ws, rank, world)
            n_recv = int(info.recv_rank_count_cumsum[-1])
            here = (lids[:n_recv] >= rank * epr) & (lids[:n_recv] < (rank + 1) * epr)
            coef = torch.where(here, (lids[:n_recv] + 1).float() * lsc[:n_recv], torch.zeros_like(lsc[:n_recv])).sum(-1)
            y = torch.zeros_like(recv)
            y[:n_recv] = recv[:n_recv] * coef[:, None]
            out = ta.MnnvlMoe.mnnvl_moe_alltoallv_combine(y, info, ws, rank, world, K, T)
            ref = x * ((ids + 1).float() * sc).sum(-1)[:, None]          # every expert e scales by (e + 1) * routing weight
            worst = max(worst, float((out - ref).abs().max() / ref.abs().max()))
        errs[rank] = worst
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_legacy_alltoall_round_trip_gloo(world):
    import torch.multiprocessing as mp

    errs = mp.Manager().dict()
    mp.spawn(_gloo_worker, args=(world, _free_port(), errs), nprocs=world, join=True)
    assert len(errs) == world and max(errs.values()) < 1e-5, dict(errs)
