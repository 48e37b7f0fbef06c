"""Run one MoE forward (for ncu launch listings)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flashinfer_b200.fused_moe import moe_forward, route
T, E, K, H, I = [int(x) for x in (sys.argv[1:6] if len(sys.argv) > 5 else (4096, 32, 8, 7168, 2048))]
x = (torch.randn(T, H, device="cuda") * 0.5).bfloat16()
w1 = (torch.randn(E, 2 * I, H, device="cuda") / H ** 0.5).bfloat16()
w2 = (torch.randn(E, H, I, device="cuda") / I ** 0.5).bfloat16()
logits = torch.randn(T, E, device="cuda")
for _ in range(3):
    ids, w = route(logits, None, K, 1)
    moe_forward(x, ids, w, w1, w2)
torch.cuda.synchronize()
