"""Intra-kernel timeline of decode_linear (FIB200_ENABLE_PROFILER build): per CTA, per warp role (TMA producer / MMA issuer /
epilogue): setup, wait for the previous grid (PDL), main loop, split-K exchange, epilogue.  Writes a Perfetto JSON per shape
and prints a summary table (median over CTAs, ns).  Usage: python tools/profile_decode_linear.py [outdir]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from flashinfer_b200 import profiler
from flashinfer_b200.gemm import decode_linear as dl

out_dir = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
os.makedirs(out_dir, exist_ok=True)
MAXE = 32
rows = []
for name, n, k, epi in (("o_proj", 4096, 4096, dl.EPI_RESIDUAL), ("down_proj", 4096, 14336, dl.EPI_RESIDUAL), ("gate_up", 28672, 4096, dl.EPI_GATED_SILU)):
    x = torch.randn(64, k, device="cuda", dtype=torch.bfloat16)
    w = dl.to_block_major_k(torch.randn(n, k, device="cuda", dtype=torch.bfloat16) * 0.02)
    res = torch.zeros(64, n, device="cuda", dtype=torch.bfloat16)
    sq = torch.zeros(64, device="cuda")
    ss = torch.rand(64, device="cuda") * k + 1
    kw = dict(residual=res, sumsq_out=sq) if epi == dl.EPI_RESIDUAL else dict(row_sumsq=ss)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        dl.decode_linear(x, w, epi, **kw)
    flush.zero_()  # cold L2 for the profiled launch
    torch.cuda.synchronize()
    buf = profiler.alloc_profiler_buffer(148, 3, MAXE)
    with dl.profiled(buf, MAXE):
        dl.decode_linear(x, w, epi, **kw)
    torch.cuda.synchronize()
    path = os.path.join(out_dir, f"decode_linear_{name}.perfetto.json")
    profiler.export_to_perfetto_trace(buf, dl.PROFILER_EVENTS, path, MAXE, dl.PROFILER_GROUPS)
    ev = profiler.decode_profiler_buffer(buf, MAXE)
    t0 = min(e["t_ns"] for e in ev)
    t1 = max(e["t_ns"] for e in ev)
    durs = {}
    open_ = {}
    for e in ev:
        key = (e["block"], e["group"], e["event"])
        if e["type"] == 0:
            open_[key] = e["t_ns"]
        elif e["type"] == 1 and key in open_:
            durs.setdefault((e["group"], e["event"]), []).append(e["t_ns"] - open_.pop(key))
    summ = {f"{dl.PROFILER_GROUPS[g]}.{dl.PROFILER_EVENTS[evn]}": int(sorted(v)[len(v) // 2]) for (g, evn), v in sorted(durs.items())}
    rows.append({"shape": name, "n": n, "k": k, "ctas": len({e["block"] for e in ev}), "span_ns": int(t1 - t0), "median_ns": summ})
    print(name, rows[-1], flush=True)
print("RESULT_JSON", json.dumps(rows))
