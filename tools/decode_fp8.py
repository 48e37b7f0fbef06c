"""fp8-KV decode on the tcgen05 kernel: accuracy vs the de-quantised reference and speed vs bf16 KV (B=64, kv=4096)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flashinfer_b200 as fi
from flashinfer_b200 import reference
from flashinfer_b200.testing import bench_gpu_time
def med(t):
    t = sorted(t); return t[len(t)//2]
B, kv, hq, hkv, d, ps = 64, 4096, 32, 8, 128, 16
torch.manual_seed(0)
npg = kv // ps
kc = torch.randn(B * npg, ps, hkv, d, device="cuda", dtype=torch.bfloat16) * 0.5
vc = torch.randn(B * npg, ps, hkv, d, device="cuda", dtype=torch.bfloat16) * 0.5
q = torch.randn(B, hq, d, device="cuda", dtype=torch.bfloat16)
indptr = torch.arange(0, (B + 1) * npg, npg, dtype=torch.int32); indices = torch.randperm(B * npg).int(); last = torch.full((B,), ps, dtype=torch.int32)
for name, dt in [("bf16", torch.bfloat16), ("e4m3", torch.float8_e4m3fn), ("e5m2", torch.float8_e5m2)]:
    k8, v8 = kc.to(dt), vc.to(dt)
    w = fi.BatchDecodeWithPagedKVCacheWrapper(torch.empty(128 << 20, dtype=torch.uint8, device="cuda"), "NHD")
    w.plan(indptr, indices, last, hq, hkv, d, ps, q_data_type=torch.bfloat16, kv_data_type=dt)
    out = w.run(q, (k8, v8))
    ref, _ = reference.batch_paged_attention_ref(q[:4], torch.arange(5, dtype=torch.int32), k8.float().bfloat16(), v8.float().bfloat16(), indptr[:5], indices.cpu(), last[:4], "NHD", True, 1 / math.sqrt(d), 0.0, -1)
    err = (out[:4].float() - ref.float()).abs().max().item()
    t = med(bench_gpu_time(lambda: w.run(q, (k8, v8), out=out), use_cuda_graph=True, num_iters_within_graph=5, dry_run_iters=3, repeat_iters=20))
    byts = B * kv * hkv * d * 2 * k8.element_size()
    print(f"{name}: {t*1e3:.1f} us  {byts/t/1e9:.2f} TB/s  max_err {err:.4f} (ref max {ref.abs().max().item():.3f})", flush=True)
