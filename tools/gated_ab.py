"""A/B in one process: gate_up GEMM + silu_and_mul vs the gated-epilogue GEMM (Llama-3-8B, M=64, rotating cold weights)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flashinfer_b200 as fi
from flashinfer_b200.gemm import interleave_gate_up, linear, linear_gated_silu

M, N2, K = 64, 28672, 4096
nbuf = 3
ws = [torch.randn(N2, K, device="cuda", dtype=torch.bfloat16) * 0.02 for _ in range(nbuf)]
wi = [interleave_gate_up(w) for w in ws]
x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
gu = torch.empty(M, N2, device="cuda", dtype=torch.bfloat16)
act = torch.empty(M, N2 // 2, device="cuda", dtype=torch.bfloat16)
wd = torch.randn(4096, N2 // 2, device="cuda", dtype=torch.bfloat16) * 0.02
y = torch.empty(M, 4096, device="cuda", dtype=torch.bfloat16)

def plain(i):
    linear(x, ws[i], out=gu); fi.silu_and_mul(gu, out=act); linear(act, wd, out=y)
def gated(i):
    linear_gated_silu(x, wi[i], out=act); linear(act, wd, out=y)
res = {}
for name, fn in (("plain", plain), ("gated", gated), ("plain2", plain), ("gated2", gated)):
    for i in range(nbuf): fn(i)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(nbuf): fn(i)
    torch.cuda.synchronize()
    ts = []
    for _ in range(9):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / nbuf)
    res[name] = sorted(ts)[len(ts) // 2]
print({k: round(v, 2) for k, v in res.items()})
