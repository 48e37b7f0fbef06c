"""Large dense bf16 GEMMs: linear() through the CTA-pair kernel (cta_group::2) vs the 1-CTA persistent kernel vs cuBLAS."""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

shapes = [(8192, 8192, 8192), (4096, 8192, 4096), (16384, 8192, 1024), (32768, 8192, 3584), (4096, 14336, 4096), (2048, 4096, 4096)]


def run(label):
    from flashinfer_b200.gemm.dense import linear
    res = {}
    for m, n, k in shapes:
        a = torch.randn(m, k, device="cuda", dtype=torch.bfloat16)
        w = torch.randn(n, k, device="cuda", dtype=torch.bfloat16) * 0.05
        out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        fns = {label: lambda: linear(a, w, out=out)}
        if label == "pair":
            fns["cublas"] = lambda: torch.matmul(a, w.t(), out=out)
            ref = (a[:64].float() @ w.float().t())
            linear(a, w, out=out)
            err = float((out[:64].float() - ref).abs().max() / ref.abs().max())
            res[f"{m}x{n}x{k}_relerr"] = round(err, 5)
        for name, fn in fns.items():
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = sorted(ts)[len(ts) // 2]
            res[f"{m}x{n}x{k}_{name}_tflops"] = round(2.0 * m * n * k / t / 1e9, 1)
    print("RESULT_JSON", json.dumps(res), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
    else:
        for label, env in (("pair", "1"), ("one_cta", "0")):
            subprocess.run([sys.executable, __file__, label], env=dict(os.environ, FIB200_GEMM_2CTA=env))
