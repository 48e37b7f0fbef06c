"""FIB200_GEMM_TRACE=1: clock64 stamps of the cluster split-K GEMM path (M=64 decode shapes)."""
import os, sys
os.environ["FIB200_GEMM_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flashinfer_b200 as fi
from flashinfer_b200.gemm import dense

N, K = (int(x) for x in (sys.argv[1:3] or ("4096", "4096")))
ws = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02 for _ in range(8)]
a = torch.randn(64, K, device="cuda", dtype=torch.bfloat16)
for w in ws[:6]:
    fi.mm_bf16(a, w.t())
torch.cuda.synchronize()
wsp = dense._workspace(a.device)
wsp.zero_()
torch.cuda.synchronize()
fi.mm_bf16(a, ws[7].t())  # cold weights
torch.cuda.synchronize()
tr = wsp.view(torch.uint8)[4096:4096 + 148 * 64].view(torch.int64).view(-1, 8).cpu()
tr = tr[tr[:, 7] != 0]
print("ctas traced:", tr.shape[0])
names = ["setup_done", "first_full", "kb8_full", "last_full", "acc_done", "epi_done", "exit"]
rel = tr[:, :7] - tr[:, 7:8]
for i, n in enumerate(names):
    c = rel[:, i]
    c = c[c > 0]
    if c.numel():
        print(f"{n:12s} min {int(c.min()):7d}  median {int(c.median()):7d}  max {int(c.max()):7d}")
