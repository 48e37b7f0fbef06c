"""One-shot push all-reduce with fused prologue / epilogue (csrc/comm/allreduce.cu: allreduce_push_kernel) through the
reference-named API: every AllReduceFusionPattern (raw sum, + residual + RMSNorm, fp8 / NVFP4 quantisation with the three
scale-factor layouts), allreduce_out, the MoE reduction / MoE finalize fusions, CUDA-graph replay - against NCCL + torch
oracles.  Port of reference tests/comm/test_trtllm_allreduce_fusion.py:27-110, test_trtllm_moe_allreduce_fusion*.py."""
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


E2M1 = [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0]


def _dequant_fp4(q, sf, layout, rows, hidden, gs):
    """packed e2m1 [rows, hidden/2] + scale bytes in `layout` -> float [rows, hidden] (test-side inverse of the epilogue)."""
    lut = torch.tensor(E2M1 + [-v for v in E2M1], device=q.device)
    b = q.view(torch.uint8).reshape(rows, hidden // 2)
    vals = torch.stack([lut[(b & 0xF).long()], lut[(b >> 4).long()]], -1).reshape(rows, hidden)
    ncol = hidden // 16
    pad4 = (ncol + 3) // 4 * 4
    r = torch.arange(rows, device=q.device)[:, None]
    c = torch.arange(ncol, device=q.device)[None, :]
    if layout == 2:
        off = r * ncol + c
    elif layout == 1:
        off = ((r // 8) * (pad4 // 4) + c // 4) * 32 + (r % 8) * 4 + (c % 4)
    else:
        off = ((r // 128) * (pad4 // 4) + c // 4) * 512 + (r % 32) * 16 + ((r % 128) // 32) * 4 + (c % 4)
    s = sf.view(torch.uint8).reshape(-1)[off.reshape(-1)].view(torch.float8_e4m3fn).float().reshape(rows, ncol)
    return vals * s.repeat_interleave(16, 1) / gs


def _worker(rank, world, port, errs):
    import torch.distributed as dist

    import flashinfer_b200.comm as comm

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        P, L = comm.AllReduceFusionPattern, comm.QuantizationSFLayout
        worst = {}
        hidden, eps = 4096, 1e-5
        _, ws = comm.trtllm_create_ipc_workspace_for_all_reduce_fusion(rank, world, 256, hidden, group=dist.group.WORLD)
        it = 0
        for tokens in (1, 17, 128, 64, 200):
            for pattern, layout in ((P.kAllReduce, None), (P.kARResidualRMSNorm, None), (P.kARResidualRMSNormFP8Quant, None),
                                    (P.kARResidualRMSNormOutFP8Quant, None), (P.kARResidualRMSNormFP4Quant, L.SWIZZLED_128x4),
                                    (P.kARResidualRMSNormOutFP4Quant, L.LINEAR), (P.kARResidualRMSNormFP4Quant, L.SWIZZLED_8x4)):
                it += 1
                torch.manual_seed(1000 * it + rank)
                x = torch.randn(tokens, hidden, device="cuda").bfloat16()
                torch.manual_seed(77 * it)
                res = torch.randn(tokens, hidden, device="cuda").bfloat16()
                gamma = (1 + 0.1 * torch.randn(hidden, device="cuda")).bfloat16()
                tot = x.float().clone()
                dist.all_reduce(tot)
                ar_out = torch.empty_like(x)
                res_out, norm_out = torch.empty_like(x), torch.empty_like(x)
                q8 = torch.empty(tokens, hidden, dtype=torch.float8_e4m3fn, device="cuda")
                q4 = torch.empty(tokens, hidden // 2, dtype=torch.uint8, device="cuda")
                sfo = torch.zeros(comm.compute_fp4_swizzled_layout_sf_size(tokens, hidden // 16) + 64, dtype=torch.uint8, device="cuda")
                fp4 = pattern in (P.kARResidualRMSNormFP4Quant, P.kARResidualRMSNormOutFP4Quant)
                fp8 = pattern in (P.kARResidualRMSNormFP8Quant, P.kARResidualRMSNormOutFP8Quant)
                sfac = torch.tensor([0.05 if fp8 else 448.0 * 6.0 / 8.0], device="cuda")
                comm.trtllm_allreduce_fusion(x, world, rank, tokens, hidden, ws, True, True, False, pattern, None, ar_out,
                                             None if pattern == P.kAllReduce else res, None if pattern == P.kAllReduce else res_out,
                                             None if pattern == P.kAllReduce else norm_out, q8 if fp8 else (q4 if fp4 else None),
                                             sfo if fp4 else None, None if pattern == P.kAllReduce else gamma, eps,
                                             sfac if (fp4 or fp8) else None, layout)
                torch.cuda.synchronize()
                key = f"p{pattern}"
                e = float((ar_out.float() - tot).abs().max() / tot.abs().max())
                if pattern != P.kAllReduce:
                    r_ref = (tot.bfloat16().float() + res.float())
                    n_ref = r_ref * torch.rsqrt(r_ref.pow(2).mean(-1, keepdim=True) + eps) * gamma.float()
                    e = max(e, float((res_out.float() - r_ref).abs().max() / r_ref.abs().max()))
                    e = max(e, float((norm_out.float() - n_ref).abs().max() / n_ref.abs().max()))
                    if fp8:
                        e = max(e, float((q8.float() * 0.05 - norm_out.float()).abs().max() / n_ref.abs().max()) / 4)  # e4m3: 2^-3 rel
                    if fp4:
                        deq = _dequant_fp4(q4, sfo, layout, tokens, hidden, float(sfac))
                        cos = torch.nn.functional.cosine_similarity(deq.flatten(), norm_out.float().flatten(), dim=0)
                        e = max(e, float(1 - cos) * 0.3)  # nvfp4 round trip: cosine > 0.97 -> < 0.01
                worst[key] = max(worst.get(key, 0.0), e)
        # ---- MoE reduction + AR + residual + RMSNorm in one kernel
        E, T = 4, 33
        torch.manual_seed(5 + rank)
        act = torch.randn(E, T, hidden, device="cuda").bfloat16()
        sc = torch.rand(E, T, device="cuda")
        tok = torch.randn(T, hidden, device="cuda").bfloat16()
        torch.manual_seed(6)
        res = torch.randn(T, hidden, device="cuda").bfloat16()
        gamma = (1 + 0.1 * torch.randn(hidden, device="cuda")).bfloat16()
        local = ((act.float() * sc[..., None]).sum(0) + tok.float()).bfloat16().float()
        tot = local.clone()
        dist.all_reduce(tot)
        res_out, norm_out, ar_o = torch.empty_like(res), torch.empty_like(res), torch.empty_like(res)
        comm.trtllm_moe_allreduce_fusion(world, rank, T, hidden, ws, True, res, gamma, eps, None, E, sc, act, tok, moe_allreduce_out=ar_o,
                                         residual_out=res_out, norm_out=norm_out)
        torch.cuda.synchronize()
        r_ref = tot.bfloat16().float() + res.float()
        n_ref = r_ref * torch.rsqrt(r_ref.pow(2).mean(-1, keepdim=True) + eps) * gamma.float()
        worst["moe_reduction"] = max(float((res_out.float() - r_ref).abs().max() / r_ref.abs().max()),
                                     float((norm_out.float() - n_ref).abs().max() / n_ref.abs().max()),
                                     float((ar_o.float() - tot).abs().max() / tot.abs().max()))
        # ---- MoE finalize + AR + residual + RMSNorm in one kernel
        K, Pn = 3, 150
        torch.manual_seed(9 + rank)
        rows = torch.randn(Pn, hidden, device="cuda").bfloat16()
        idx = torch.randint(0, Pn, (T, K), device="cuda", dtype=torch.int32)
        wts = torch.rand(T, K, device="cuda")
        shared = torch.randn(T, hidden, device="cuda").bfloat16()
        local = ((rows[idx.long()].float() * wts[..., None]).sum(1) + shared.float()).bfloat16().float()
        tot = local.clone()
        dist.all_reduce(tot)
        comm.trtllm_moe_finalize_allreduce_fusion(rows, res, gamma, idx, norm_out, res_out, None, None, ws, True, rank, world, eps, shared, wts)
        torch.cuda.synchronize()
        r_ref = tot.bfloat16().float() + res.float()
        n_ref = r_ref * torch.rsqrt(r_ref.pow(2).mean(-1, keepdim=True) + eps) * gamma.float()
        worst["moe_finalize"] = max(float((res_out.float() - r_ref).abs().max() / r_ref.abs().max()),
                                    float((norm_out.float() - n_ref).abs().max() / n_ref.abs().max()))
        # ---- CUDA-graph replay (device-side epoch / dirty-row bookkeeping)
        x = torch.randn(64, hidden, device="cuda").bfloat16()
        out = torch.empty_like(x)
        s_ = torch.cuda.Stream()
        with torch.cuda.stream(s_):
            comm.allreduce_fusion(x, ws, P.kAllReduce, True, output=out)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(5):
                comm.allreduce_fusion(x, ws, P.kAllReduce, True, output=out)
        for _ in range(4):
            g.replay()
        torch.cuda.synchronize()
        tot = x.float().clone()
        dist.all_reduce(tot)
        worst["graph"] = float((out.float() - tot).abs().max() / tot.abs().max())
        errs[rank] = worst
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_allreduce_push_patterns(world):
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    errs = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), errs), nprocs=world, join=True)
    assert len(errs) == world
    for r, w in errs.items():
        for k, v in w.items():
            assert v < 2e-2, (r, k, v)
