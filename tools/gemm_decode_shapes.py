"""Llama-3-8B decode GEMM shapes at M=64 (weights streamed from HBM, rotating buffers > L2): ours vs cuBLAS vs the HBM floor."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flashinfer_b200 as fi

M = int(os.environ.get("M", "64"))
shapes = {"qkv": (6144, 4096), "o": (4096, 4096), "gate_up": (28672, 4096), "down": (4096, 14336), "lm_head": (128256, 4096)}
peak = 6.57e12
res = {}
only = os.environ.get("SHAPES")
if only:
    shapes = {k: v for k, v in shapes.items() if k in only.split(",")}
nocublas = os.environ.get("NOCUBLAS") == "1"
for name, (N, K) in shapes.items():
    nbuf = max(2, int(300e6 // (N * K * 2)) + 1)
    ws = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02 for _ in range(nbuf)]
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    outs = {}
    for label, fn in (("ours", lambda w: fi.mm_bf16(a, w.t())), ("cublas", lambda w: a @ w.t())):
        if nocublas and label == "cublas":
            outs[label] = 0.0
            continue
        for i in range(nbuf):
            fn(ws[i])
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(nbuf):
                fn(ws[i])
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / nbuf)
        outs[label] = sorted(ts)[len(ts) // 2]
    floor = N * K * 2 / peak * 1e6
    res[name] = {"ours_us": round(outs["ours"], 2), "cublas_us": round(outs["cublas"], 2), "floor_us": round(floor, 2),
                 "ours_frac_of_floor": round(floor / outs["ours"], 3)}
    print(name, res[name], flush=True)
    del ws
print("RESULT_JSON", json.dumps(res))
