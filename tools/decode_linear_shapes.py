"""Decode-step projections at M = 64 (rotating weight buffers > L2, CUDA-graph timed back to back, PDL chained):
decode_linear (new fused-epilogue kernel) vs gemm_nt (round-1 kernel) vs cuBLAS, Llama-3-8B shapes at TP 1 / 8.
Env: TP=1|8, VARIANTS="bn:s:kb:layout,..." extra (BN, split-K, ring KB, 1 = BlockMajorK weights) variants of decode_linear to sweep."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import flashinfer_b200 as fi
from flashinfer_b200.gemm import decode_linear as dl
from flashinfer_b200.gemm.dense import linear, linear_gated_silu

M = int(os.environ.get("M", "64"))
TP = int(os.environ.get("TP", "1"))
H, I, HQ, HKV, D = 4096, 14336, 32, 8, 128
shapes = {"qkv": ((HQ + 2 * HKV) * D // TP, H, "rope"), "o": (H, HQ * D // TP, "resid"), "gate_up": (2 * I // TP, H, "gated"),
          "down": (H, I // TP, "resid"), "lm_head": ((128256 + TP - 1) // TP // 16 * 16, H, "plain")}
only = os.environ.get("SHAPES")
if only:
    shapes = {k: v for k, v in shapes.items() if k in only.split(",")}
variants = [(0, 0, 0, 0)] + [tuple(int(t) for t in (v.split(":") + ["0"])[:4]) for v in os.environ.get("VARIANTS", "").split(",") if v]
peak = 6.57e12
res = {}


def timed(fn, nbuf):
    for i in range(nbuf):
        fn(i)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(nbuf):
            fn(i)
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / nbuf)
    return sorted(ts)[len(ts) // 2]


for name, (N, K, kind) in shapes.items():
    nbuf = max(4, int(400e6 // (N * K * 2)) + 1)
    ws = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02 for _ in range(nbuf)]
    wbs = None
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    ss = torch.rand(64, device="cuda") * K + 1
    out = {}
    if kind == "rope":
        hq, hkv = HQ // TP, HKV // TP
        cs = torch.rand(64, D, device="cuda")
        kc = torch.zeros(64, 16, hkv, D, device="cuda", dtype=torch.bfloat16)
        vc = torch.zeros_like(kc)
        rows = (torch.arange(64, device="cuda") * 16 * hkv * D).long()
        q = torch.empty(M, hq * D, device="cuda", dtype=torch.bfloat16)
        mk = lambda bn, s, kb, W: (lambda i: dl.decode_linear(a, W[i], dl.EPI_ROPE_APPEND, out=q, row_sumsq=ss, cos_sin=cs, cache_row=rows,
                                                           k_cache=kc, v_cache=vc, num_q_heads=hq, num_kv_heads=hkv, head_dim=D, bn=bn,
                                                           split_k=s, smem_kb=kb))
        o1 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        old = lambda i: linear(a, ws[i], out=o1)
    elif kind == "gated":
        o2 = torch.empty(M, N // 2, device="cuda", dtype=torch.bfloat16)
        mk = lambda bn, s, kb, W: (lambda i: dl.decode_linear(a, W[i], dl.EPI_GATED_SILU, out=o2, row_sumsq=ss, bn=bn, split_k=s, smem_kb=kb))
        old = lambda i: linear_gated_silu(a, ws[i], out=o2)
    elif kind == "plain":
        o1 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        mk = lambda bn, s, kb, W: (lambda i: dl.decode_linear(a, W[i], dl.EPI_PLAIN, out=o1, row_sumsq=ss, bn=bn, split_k=s, smem_kb=kb))
        old = lambda i: linear(a, ws[i], out=o1)
    else:
        r = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        sq = torch.zeros(64, device="cuda")
        mk = lambda bn, s, kb, W: (lambda i: dl.decode_linear(a, W[i], dl.EPI_RESIDUAL, residual=r, sumsq_out=sq, bn=bn, split_k=s, smem_kb=kb))
        o1 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        old = lambda i: linear(a, ws[i], out=o1)
    for bn, s, kb, lay in variants:
        key = f"dl[{bn}:{s}:{kb}:{lay}]"
        try:
            if lay and wbs is None:
                wbs = [dl.to_block_major_k(w) for w in ws]
            out[key] = round(timed(mk(bn, s, kb, wbs if lay else ws), nbuf), 2)
        except Exception as e:  # noqa: BLE001
            out[key] = f"ERR {str(e)[:80]}"
    out["gemm_nt"] = round(timed(old, nbuf), 2)
    if os.environ.get("NOCUBLAS") != "1":
        out["cublas"] = round(timed(lambda i: torch.matmul(a, ws[i].t()), nbuf), 2)
    out["floor_us"] = round(N * K * 2 / peak * 1e6, 2)
    res[name] = out
    print(name, (N, K), out, flush=True)
    del ws, wbs
print("RESULT_JSON", json.dumps({"tp": TP, "m": M, "shapes": res}))
