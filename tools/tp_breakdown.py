"""Per-kernel breakdown of the fused Llama-3-8B decode step at TP = T (B = 64, kv = 4096): every kernel of a layer timed on its
own as a CUDA-graph chain over rotating buffers (cold L2), on this rank's TP-shaped tensors.  Run single-process for the
communication-free kernels (TP=8 python tools/tp_breakdown.py) or under torchrun (world == TP) to add the all-reduce variants:
  o / down GEMM   : local only | with the in-kernel all-reduce (decode_linear EPI_RESIDUAL, tp=...) | plain GEMM + the separate
                    all-reduce + add + RMSNorm kernel of round 1
  all-reduce alone: TPCommunicator.allreduce_add_rmsnorm on a [64, 4096] bf16 message (latency vs tokens: 1, 16, 64)
Output: one JSON line (rank 0)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import flashinfer_b200 as fi
from flashinfer_b200.gemm import decode_linear as dl

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
TP = int(os.environ.get("TP", str(world)))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
B, KV, PAGE, H, I, HQ, HKV, D, V = 64, 4096, 16, 4096, 14336, 32, 8, 128, 128256
hq, hkv, inter = HQ // TP, HKV // TP, I // TP
dev = "cuda"
res = {}


def timed(fn, n):
    for i in range(n):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn(i)
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / n)
    t = torch.tensor([sorted(ts)[len(ts) // 2]], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t), 2)


def weights(n, k, cnt):
    return [dl.to_block_major_k(torch.randn(n, k, device=dev, dtype=torch.bfloat16) * 0.02) for _ in range(cnt)]


NB = 8
ss = torch.rand(64, device=dev) * H + 1
x = torch.randn(B, H, device=dev, dtype=torch.bfloat16)
# ---- attention
ppr = KV // PAGE
indptr = torch.arange(0, (B + 1) * ppr, ppr, dtype=torch.int32)
indices = torch.randperm(B * ppr).int()
last = torch.full((B,), PAGE, dtype=torch.int32)
ws = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
wr = fi.BatchDecodeWithPagedKVCacheWrapper(ws, "NHD")
wr.plan(indptr, indices, last, hq, hkv, D, PAGE, q_data_type=torch.bfloat16)
caches = [(torch.randn(B * ppr, PAGE, hkv, D, device=dev, dtype=torch.bfloat16), torch.randn(B * ppr, PAGE, hkv, D, device=dev, dtype=torch.bfloat16))
          for _ in range(4)]
q = torch.randn(B, hq, D, device=dev, dtype=torch.bfloat16)
ao = torch.empty_like(q)
res["attention"] = timed(lambda i: wr.run(q, caches[i % 4], out=ao), 8)
res["attention_floor"] = round(B * KV * hkv * D * 2 * 2 / 6.57e12 * 1e6, 2)
del caches
# ---- qkv (+rope+append), gate/up
nqkv = (hq + 2 * hkv) * D
w = weights(nqkv, H, NB)
cs = torch.rand(64, D, device=dev)
kc = torch.zeros(64, 16, hkv, D, device=dev, dtype=torch.bfloat16)
vc = torch.zeros_like(kc)
rows = (torch.arange(64, device=dev) * 16 * hkv * D).long()
qo = torch.empty(B, hq * D, device=dev, dtype=torch.bfloat16)
res["qkv"] = timed(lambda i: dl.decode_linear(x, w[i % NB], dl.EPI_ROPE_APPEND, out=qo, row_sumsq=ss, cos_sin=cs, cache_row=rows, k_cache=kc,
                                              v_cache=vc, num_q_heads=hq, num_kv_heads=hkv, head_dim=D), 16)
res["qkv_floor"] = round(nqkv * H * 2 / 6.57e12 * 1e6, 2)
w = weights(2 * inter, H, NB)
act = torch.empty(B, inter, device=dev, dtype=torch.bfloat16)
res["gate_up"] = timed(lambda i: dl.decode_linear(x, w[i % NB], dl.EPI_GATED_SILU, out=act, row_sumsq=ss), 16)
res["gate_up_floor"] = round(2 * inter * H * 2 / 6.57e12 * 1e6, 2)
vs = (V + TP - 1) // TP
w = weights(vs // 16 * 16, H, 2)
lo = torch.empty(B, vs // 16 * 16, device=dev, dtype=torch.bfloat16)
res["lm_head"] = timed(lambda i: dl.decode_linear(x, w[i % 2], out=lo, row_sumsq=ss), 4)
res["lm_head_floor"] = round(vs * H * 2 / 6.57e12 * 1e6, 2)
del w
# ---- o / down
tp = dl.FusedLinearTP(None, 64, H, algo=1) if world > 1 else None
tp2 = dl.FusedLinearTP(None, 64, H, algo=2) if world > 1 and 64 % world == 0 else None
comm = None
if world > 1:
    from flashinfer_b200.comm import TPCommunicator

    comm = TPCommunicator(dist.group.WORLD, max_tokens=B, hidden=H, dtype=torch.bfloat16)
resid = torch.zeros(B, H, device=dev, dtype=torch.bfloat16)
sq = torch.zeros(64, device=dev)
gamma = torch.ones(H, device=dev, dtype=torch.bfloat16)
for name, k in (("o", hq * D), ("down", inter)):
    w = weights(H, k, NB)
    a = torch.randn(B, k, device=dev, dtype=torch.bfloat16)
    res[f"{name}_local"] = timed(lambda i: dl.decode_linear(a, w[i % NB], dl.EPI_RESIDUAL, residual=resid, sumsq_out=sq), 16)
    res[f"{name}_floor"] = round(H * k * 2 / 6.57e12 * 1e6, 2)
    if world > 1:
        res[f"{name}_fused_ar_one_shot"] = timed(lambda i: dl.decode_linear(a, w[i % NB], dl.EPI_RESIDUAL, residual=resid, sumsq_out=sq, tp=tp), 16)
        if tp2 is not None:
            res[f"{name}_fused_ar_two_shot"] = timed(lambda i: dl.decode_linear(a, w[i % NB], dl.EPI_RESIDUAL, residual=resid, sumsq_out=sq, tp=tp2), 16)
        xo = torch.empty(B, H, device=dev, dtype=torch.bfloat16)

        def sep(i):
            part = dl.decode_linear(a, w[i % NB], out=comm.gemm_out(B))
            comm.allreduce_add_rmsnorm(part, resid, gamma, 1e-5, out=xo)

        res[f"{name}_gemm_plus_ar_kernel"] = timed(sep, 16)
    del w
if world > 1:
    xo = torch.empty(B, H, device=dev, dtype=torch.bfloat16)
    for toks in (1, 16, 64):
        def ar(i, toks=toks):
            part = comm.gemm_out(toks)
            comm.allreduce_add_rmsnorm(part, resid[:toks], gamma, 1e-5, out=xo[:toks])

        res[f"ar_kernel_alone_{toks}tok"] = timed(ar, 16)
    buf = torch.randn(B, H, device=dev, dtype=torch.bfloat16)
    res["nccl_allreduce_64tok"] = timed(lambda i: dist.all_reduce(buf), 16)
if rank == 0:
    print("RESULT_JSON", json.dumps({"tp": TP, "world": world, "us": res}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
