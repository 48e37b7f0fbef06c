"""Does the symmetric heap get an NVLS multicast alias at this size / world?  torchrun --nproc-per-node N tools/mc_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
from flashinfer_b200.comm.symm import SymmetricHeap
for mb in (64, 512, 1100, 1600, 2200):
    try:
        h = SymmetricHeap(None, mb << 20)
        if dist.get_rank() == 0:
            print(f"MCPROBE world={dist.get_world_size()} heap={mb}MB mc_ptr={'yes' if h.mc_ptr else 'NO'}", flush=True)
        del h
    except Exception as e:  # noqa: BLE001
        if dist.get_rank() == 0:
            print(f"MCPROBE heap={mb}MB failed: {type(e).__name__}: {str(e)[:150]}", flush=True)
dist.barrier()
dist.destroy_process_group()
