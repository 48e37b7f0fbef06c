"""mm_fp4 small-shape experiments (published shape m=512 n=1024 k=7168)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flashinfer_b200 as fi
from flashinfer_b200.testing import bench_gpu_time
def med(t):
    t = sorted(t); return t[len(t)//2]
m, n, k = 512, 1024, 7168
a = torch.randn(m, k, device="cuda", dtype=torch.bfloat16); w = torch.randn(n, k, device="cuda", dtype=torch.bfloat16)
g = torch.tensor(1.0, device="cuda")
aq, asf = fi.nvfp4_quantize(a, g); wq, wsf = fi.nvfp4_quantize(w, g)
out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
for split in ["1", "2"]:
    for bn in ["64", "128", "192", "256"]:
        os.environ["FIB200_LOWP_SPLIT"] = split; os.environ["FIB200_LOWP_BN"] = bn
        try:
            t = med(bench_gpu_time(lambda: fi.mm_fp4(aq, wq.t(), asf, wsf, g, torch.bfloat16, out), use_cuda_graph=True, num_iters_within_graph=5, dry_run_iters=3, repeat_iters=20))
            t2 = med(bench_gpu_time(lambda: fi.mm_fp4(aq, wq.t(), asf, wsf, g, torch.bfloat16, out), use_cuda_graph=True, num_iters_within_graph=5, dry_run_iters=3, repeat_iters=20, l2_flush=False))
            print(f"split {split} bn {bn}: cold {t*1e3:.2f} us warm {t2*1e3:.2f} us", flush=True)
        except Exception as e:
            print(split, bn, repr(e)[:150])
