"""One NVFP4 MoE forward on the published shape (for ncu launch listings)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flashinfer_b200.fused_moe import RoutingMethodType, trtllm_fp4_block_scale_moe
from flashinfer_b200.quantization.fp4 import fp4_quantize
T, H, I, E, K = 1024, 1024, 1024, 256, 8
x = torch.randn(T, H, device="cuda", dtype=torch.bfloat16)
w1 = torch.randn(E, 2 * I, H, device="cuda", dtype=torch.bfloat16) * 0.03
w2 = torch.randn(E, H, I, device="cuda", dtype=torch.bfloat16) * 0.03
logits = torch.randn(T, E, device="cuda"); bias = torch.randn(E, device="cuda") * 0.1
def qw(wt):
    qs, sfs = [], []
    for e in range(wt.shape[0]):
        q, sf = fp4_quantize(wt[e], torch.ones(1, device="cuda"), 16, False, False)
        qs.append(q); sfs.append(sf)
    return torch.stack(qs), torch.stack(sfs)
w1q, w1sf = qw(w1); w2q, w2sf = qw(w2)
one = torch.ones(E, device="cuda")
for _ in range(3):
    trtllm_fp4_block_scale_moe(logits, bias, x, None, w1q, w1sf, None, None, None, None, w2q, w2sf, None, one, one, one, E, K, 8, 4, I, 0, E, 2.5, RoutingMethodType.DeepSeekV3)
torch.cuda.synchronize()
