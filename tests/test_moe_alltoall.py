"""EP MoE through MoeAlltoAll dispatch -> local experts -> combine == single-process MoE
(reference tests/comm/test_trtllm_moe_alltoall.py: spawned ranks, per-rank comparison with a torch oracle)."""
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


def _ep_moe_worker(rank, world, port, backend, errs):
    from flashinfer_b200.comm import Mapping, MoeAlltoAll
    from flashinfer_b200.fused_moe import moe_forward, moe_reference, route

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
        E, K, H, I = 8 * world, 4, 256, 128
        epr = E // world
        g = torch.Generator().manual_seed(1234)  # same weights everywhere
        w1 = (torch.randn(E, 2 * I, H, generator=g) / H ** 0.5).to(dtype).to(dev)
        w2 = (torch.randn(E, H, I, generator=g) / I ** 0.5).to(dtype).to(dev)
        a2a = MoeAlltoAll(Mapping(world, rank, tp_size=world, moe_ep_size=world, moe_tp_size=1), max_num_tokens=512, top_k=K,
                          num_experts=E, hidden_size=H, dtype=dtype)
        worst = 0.0
        for it, T in enumerate([1 + rank, 64, 300 + 7 * rank, 64]):
            gl = torch.Generator().manual_seed(100 * it + rank)
            x = (torch.randn(T, H, generator=gl) * 0.5).to(dtype).to(dev)
            logits = torch.randn(T, E, generator=gl).to(dev)
            ids, w = route(logits, None, K, 1)
            R = 512
            recv_x, recv_ids, recv_w = a2a.dispatch(ids, [x, ids, w], R, invalid_token_expert_id=-1, expert_id_payload_index=1)
            rx = recv_x.reshape(world * R, H)
            rids = recv_ids.reshape(world * R, K)
            rw = recv_w.reshape(world * R, K)
            payload = a2a.get_combine_payload_tensor_in_workspace(R, H, dtype).view(world * R, H)
            y = moe_forward(rx, rids, rw, w1[rank * epr:(rank + 1) * epr].contiguous(), w2[rank * epr:(rank + 1) * epr].contiguous(),
                            local_expert_offset=rank * epr, num_experts=E, out=payload if cuda else None)
            if not cuda:
                payload = y
            out = a2a.combine(payload.view(world, R, H), R, payload_in_workspace=cuda)
            ref = moe_reference(x, ids, w, w1, w2)
            worst = max(worst, float((out.float() - ref).abs().max() / max(1.0, float(ref.abs().max()))))
        errs[rank] = worst
        if cuda:
            torch.cuda.synchronize()
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_moe_alltoall_gloo():
    errs = mp.Manager().dict()
    mp.spawn(_ep_moe_worker, args=(2, _free_port(), "gloo", errs), nprocs=2, join=True)
    assert max(errs.values()) < 1e-4, dict(errs)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_moe_alltoall_nvlink(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    errs = mp.Manager().dict()
    mp.spawn(_ep_moe_worker, args=(world, _free_port(), "nccl", errs), nprocs=world, join=True)
    assert max(errs.values()) < 4e-2, dict(errs)
