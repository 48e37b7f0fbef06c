"""Small-batch paged decode experiments (published shape P3: B=16, kv<=1024, 64/8 heads)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flashinfer_b200 as fi
from flashinfer_b200.testing import bench_gpu_time

def med(t):
    t = sorted(t); return t[len(t)//2]

B, kv, hq, hkv, d, ps = 16, 1024, 64, 8, 128, 16
torch.manual_seed(0)
lens = torch.randint(1, kv + 1, (B,))
npg = (lens + ps - 1) // ps
indptr = torch.zeros(B + 1, dtype=torch.int32); indptr[1:] = npg.cumsum(0)
total = int(indptr[-1])
indices = torch.randperm(total).int()
last = ((lens - 1) % ps + 1).int()
kc = torch.randn(total, ps, hkv, d, device="cuda", dtype=torch.bfloat16)
vc = torch.randn(total, ps, hkv, d, device="cuda", dtype=torch.bfloat16)
q = torch.randn(B, hq, d, device="cuda", dtype=torch.bfloat16)
out = torch.empty_like(q)
mode = sys.argv[1] if len(sys.argv) > 1 else "all"
for name, kw, budget in [("default", {}, None), ("nosplit", {"disable_split_kv": True}, None), ("budget128", {}, 128), ("budget64", {}, 64)]:
    w = fi.BatchDecodeWithPagedKVCacheWrapper(torch.empty(128 << 20, dtype=torch.uint8, device="cuda"), "NHD")
    if budget: w._cta_budget = budget
    w.plan(indptr, indices, last, hq, hkv, d, ps, q_data_type=torch.bfloat16, **kw)
    if mode == "ncu":
        if name == "default":
            for _ in range(3): w.run(q, (kc, vc), out=out)
            torch.cuda.synchronize()
        continue
    t = med(bench_gpu_time(lambda: w.run(q, (kc, vc), out=out), use_cuda_graph=True, num_iters_within_graph=5, dry_run_iters=3, repeat_iters=20))
    t2 = med(bench_gpu_time(lambda: w.run(q, (kc, vc), out=out), use_cuda_graph=True, num_iters_within_graph=5, dry_run_iters=3, repeat_iters=20, l2_flush=False))
    print(name, "cold %.2f us warm %.2f us" % (t * 1e3, t2 * 1e3), "counts", w._plan_counts[:4].tolist(), flush=True)
