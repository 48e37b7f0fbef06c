"""Tile-plan sweep of the decode linear (csrc/gemm/decode_linear_sm100.cu) on the Llama-3-8B decode shapes at TP = 1, 2, 4, 8:
every admissible (BN, split-K cluster size S) single-wave plan is timed as a CUDA-graph chain over rotating weights (cold L2) on
ONE GPU (the tensor-parallel all-reduce is not part of this sweep: EPI_RESIDUAL runs with tp=None).  Prints one JSON line per
shape with the default plan's time and the best plans; the winners feed the planner in dlinear_run."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from flashinfer_b200.gemm import decode_linear as dl

dev = "cuda"
B, H, I, HQ, HKV, D = 64, 4096, 14336, 32, 8, 128
SMS = torch.cuda.get_device_properties(0).multi_processor_count


def timed(fn, n=16, reps=5):
    for i in range(n):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn(i)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / n)
    return round(sorted(ts)[len(ts) // 2], 2)


def plans(n, k):
    kb = k // 64
    out = []
    for s in (1, 2, 4, 8):
        if kb < s:
            continue
        for bn in range(16 * s, 257, 16 * s):
            tiles = (n + bn - 1) // bn
            lim = SMS if s < 4 else SMS * 132 // 148
            if tiles * s > lim or tiles * s < 40:
                continue
            # skip plans that leave most SMs idle unless nothing else exists
            out.append((bn, s))
    return out


only = os.environ.get("TPS", "1,2,4,8").split(",")
ss = torch.rand(64, device=dev) * H + 1
x = torch.randn(B, H, device=dev, dtype=torch.bfloat16)
NB = 6
for tp in map(int, only):
    hq, hkv, inter = HQ // tp, HKV // tp, I // tp
    shapes = {"qkv": ((hq + 2 * hkv) * D, H), "gate_up": (2 * inter, H), "o": (H, hq * D), "down": (H, inter)}
    for name, (n, k) in shapes.items():
        w = [dl.to_block_major_k(torch.randn(n, k, device=dev, dtype=torch.bfloat16) * 0.02) for _ in range(NB)]
        a = torch.randn(B, k, device=dev, dtype=torch.bfloat16)
        resid = torch.zeros(B, H, device=dev, dtype=torch.bfloat16)
        sq = torch.zeros(64, device=dev)
        if name == "qkv":
            cs = torch.rand(64, D, device=dev)
            kc = torch.zeros(64, 16, hkv, D, device=dev, dtype=torch.bfloat16)
            vc = torch.zeros_like(kc)
            rows = (torch.arange(64, device=dev) * 16 * hkv * D).long()
            qo = torch.empty(B, hq * D, device=dev, dtype=torch.bfloat16)
            run = lambda i, bn=0, s=0: dl.decode_linear(x, w[i % NB], dl.EPI_ROPE_APPEND, out=qo, row_sumsq=ss, cos_sin=cs, cache_row=rows,
                                                        k_cache=kc, v_cache=vc, num_q_heads=hq, num_kv_heads=hkv, head_dim=D, bn=bn, split_k=s)
        elif name == "gate_up":
            act = torch.empty(B, inter, device=dev, dtype=torch.bfloat16)
            run = lambda i, bn=0, s=0: dl.decode_linear(x, w[i % NB], dl.EPI_GATED_SILU, out=act, row_sumsq=ss, bn=bn, split_k=s)
        else:
            run = lambda i, bn=0, s=0: dl.decode_linear(a, w[i % NB], dl.EPI_RESIDUAL, residual=resid, sumsq_out=sq, bn=bn, split_k=s)
        base = timed(run)
        res = []
        for bn, s in plans(n, k):
            try:
                t = timed(lambda i: run(i, bn, s))
            except Exception as e:  # noqa: BLE001
                continue
            res.append((t, bn, s))
        res.sort()
        print("SWEEP", json.dumps({"tp": tp, "op": name, "n": n, "k": k, "default_us": base, "best": res[:6],
                                   "floor_us": round(n * k * 2 / 6.57e12 * 1e6, 2)}), flush=True)
        del w
