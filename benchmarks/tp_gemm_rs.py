"""Llama-3-70B tensor-parallel row-parallel linears: fused GEMM -> reduce-scatter -> add-RMSNorm (one GEMM kernel + a
scale pass) vs the NCCL composition (cuBLAS matmul, reduce_scatter_tensor, fused_add_rmsnorm).  BASELINE config 5:
batch 16 x seqlen 2k tokens, TP = world size.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
        benchmarks/tp_gemm_rs.py [--tokens 32768]

Device-timed with CUDA events, max over ranks; prints one JSON line per shape on rank 0."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=str, default=str(16 * 2048), help="comma-separated token counts")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    import flashinfer_b200 as fi
    from flashinfer_b200.comm import GemmAllReduce

    hidden, inter = 8192, 28672
    token_list = [int(t) // world * world for t in args.tokens.split(",")]
    comm = GemmAllReduce(None, max(token_list), hidden, torch.bfloat16)
    peaks = {"bf16_tflops": 1640.0, "nvlink_gbs": 900.0}
    for M, (name, k_full) in [(m, s) for m in token_list for s in (("o_proj", hidden), ("down_proj", inter))]:
        K = k_full // world
        torch.manual_seed(rank)
        a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
        w = (torch.randn(hidden, K, device="cuda") / k_full ** 0.5).bfloat16()
        res = torch.randn(M // world, hidden, device="cuda").bfloat16()
        gamma = torch.ones(hidden, device="cuda", dtype=torch.bfloat16)

        def fused():
            return comm.reduce_scatter(a, w, residual=res, rms_weight=gamma, eps=1e-5)

        shard_buf = torch.empty(M // world, hidden, device="cuda", dtype=torch.bfloat16)

        def composed():
            c = a @ w.t()
            dist.reduce_scatter_tensor(shard_buf, c)
            r = res.clone()
            fi.fused_add_rmsnorm(shard_buf, r, gamma, 1e-5)
            return shard_buf, r

        out = {}
        for label, fn in (("fused", fused), ("nccl_composed", composed)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            ts = []
            for _ in range(args.iters):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = torch.tensor(sorted(ts)[len(ts) // 2], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out[label] = float(t)
        flops = 2.0 * M * hidden * K
        t_compute = flops / (peaks["bf16_tflops"] * 1e12) * 1e3
        # reduce-scatter: every rank receives (world-1)/world of its shard sum inputs = M*N*2 * (world-1)/world bytes in
        t_link = M * hidden * 2 * (world - 1) / world / (peaks["nvlink_gbs"] * 1e9) * 1e3
        if rank == 0:
            print(json.dumps({
                "shape": f"{name} M={M} N={hidden} K_local={K} tp={world}", "fused_ms": round(out["fused"], 4),
                "nccl_composed_ms": round(out["nccl_composed"], 4), "speedup": round(out["nccl_composed"] / out["fused"], 3),
                "roofline_ms": round(max(t_compute, t_link), 4), "fraction_of_roofline": round(max(t_compute, t_link) / out["fused"], 3),
            }), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
