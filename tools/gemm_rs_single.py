"""World-size-1 run of the fused GEMM -> reduce-scatter kernel (profiling aid: same GEMM / epilogue / flag protocol, the
"reduction" pulls from the local staging buffer)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
import torch, torch.distributed as dist
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
from flashinfer_b200.comm import GemmAllReduce
import flashinfer_b200 as fi
M, N, K = (int(x) for x in (sys.argv[1:4] or ("4096", "8192", "4096")))
comm = GemmAllReduce(None, M, N, torch.bfloat16, use_nvls=(os.environ.get("NVLS", "1") == "1"))
a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16(); w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
res = torch.randn(M, N, device="cuda").bfloat16(); g = torch.ones(N, device="cuda", dtype=torch.bfloat16)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("fused rs(+norm) ms", t(lambda: comm.reduce_scatter(a, w, residual=res, rms_weight=g)))
print("fused rs        ms", t(lambda: comm.reduce_scatter(a, w)))
print("fused ar 1shot  ms", t(lambda: comm(a, w, two_shot=False)))
print("mm_bf16         ms", t(lambda: fi.mm_bf16(a, w.t())))
print("cublas          ms", t(lambda: a @ w.t()))
dist.destroy_process_group()
