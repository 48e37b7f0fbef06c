"""Single-kernel workloads for `ncu --set full` captures:  python tools/prof_targets.py <gemm_bf16|gemm_fp4|gemm_fp8|prefill|decode|moe|mla|mla_small>"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flashinfer_b200 as fi

mode = sys.argv[1]
torch.manual_seed(0)
if mode == "gemm_bf16":
    a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16); w = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    for _ in range(3): fi.mm_bf16(a, w.t())
elif mode in ("gemm_fp4", "gemm_fp8"):
    a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16); w = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    g = torch.tensor(1.0, device="cuda")
    if mode == "gemm_fp4":
        aq, asf = fi.nvfp4_quantize(a, g); wq, wsf = fi.nvfp4_quantize(w, g)
        for _ in range(3): fi.mm_fp4(aq, wq.t(), asf, wsf, g, torch.bfloat16)
    else:
        a8, w8 = a.to(torch.float8_e4m3fn), w.to(torch.float8_e4m3fn)
        for _ in range(3): fi.mm_fp8(a8, w8.t(), g)
elif mode == "prefill":
    L, hq, hkv, d = 8192, 32, 8, 128
    q = torch.randn(L, hq, d, device="cuda", dtype=torch.bfloat16); k = torch.randn(L, hkv, d, device="cuda", dtype=torch.bfloat16); v = torch.randn(L, hkv, d, device="cuda", dtype=torch.bfloat16)
    w = fi.BatchPrefillWithRaggedKVCacheWrapper(torch.empty(128 << 20, dtype=torch.uint8, device="cuda"))
    ip = torch.tensor([0, L], dtype=torch.int32)
    w.plan(ip, ip, hq, hkv, d, causal=True, q_data_type=torch.bfloat16)
    for _ in range(3): w.run(q, k, v)
elif mode == "decode":
    B, kv, hq, hkv, d, ps = 64, 4096, 32, 8, 128, 16
    npg = kv // ps
    kc = torch.randn(B * npg, ps, hkv, d, device="cuda", dtype=torch.bfloat16); vc = torch.randn(B * npg, ps, hkv, d, device="cuda", dtype=torch.bfloat16)
    q = torch.randn(B, hq, d, device="cuda", dtype=torch.bfloat16)
    w = fi.BatchDecodeWithPagedKVCacheWrapper(torch.empty(128 << 20, dtype=torch.uint8, device="cuda"), "NHD")
    w.plan(torch.arange(0, (B + 1) * npg, npg, dtype=torch.int32), torch.randperm(B * npg).int(), torch.full((B,), ps, dtype=torch.int32), hq, hkv, d, ps, q_data_type=torch.bfloat16)
    for _ in range(3): w.run(q, (kc, vc))
elif mode == "moe":
    from flashinfer_b200.fused_moe import moe_forward, route
    T, E, K, H, I = 4096, 32, 8, 7168, 2048
    x = (torch.randn(T, H, device="cuda") * 0.5).bfloat16(); w1 = (torch.randn(E, 2 * I, H, device="cuda") / H ** 0.5).bfloat16(); w2 = (torch.randn(E, H, I, device="cuda") / I ** 0.5).bfloat16()
    ids, w = route(torch.randn(T, E, device="cuda"), None, K, 1)
    for _ in range(3): moe_forward(x, ids, w, w1, w2)
elif mode in ("mla", "mla_small"):
    from flashinfer_b200.mla import BatchMLAPagedAttentionWrapper
    B, kv, H, ps = (64, 4096, 128, 64) if mode == "mla" else (16, 1024, 128, 32)
    npg = kv // ps
    ckv = torch.randn(B * npg, ps, 512, device="cuda", dtype=torch.bfloat16); kpe = torch.randn(B * npg, ps, 64, device="cuda", dtype=torch.bfloat16)
    qn = torch.randn(B, H, 512, device="cuda", dtype=torch.bfloat16); qp = torch.randn(B, H, 64, device="cuda", dtype=torch.bfloat16)
    w = BatchMLAPagedAttentionWrapper(torch.empty(256 << 20, dtype=torch.uint8, device="cuda"))
    w.plan(torch.arange(B + 1, dtype=torch.int32), torch.arange(0, (B + 1) * npg, npg, dtype=torch.int32), torch.randperm(B * npg).int(),
           torch.full((B,), kv, dtype=torch.int32), H, 512, 64, ps, True, 0.07, torch.bfloat16, torch.bfloat16)
    for _ in range(3): w.run(qn, qp, ckv, kpe)
elif mode in ("gemm_o", "gemm_qkv", "cublas_o", "cublas_qkv"):
    N, K = (4096, 4096) if mode.endswith("_o") else (6144, 4096)
    ws = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02 for _ in range(6)]
    a = torch.randn(64, K, device="cuda", dtype=torch.bfloat16)
    for w in ws:
        (fi.mm_bf16(a, w.t()) if mode.startswith("gemm") else a @ w.t())
elif mode in ("dlinear_gate_up", "dlinear_down", "dlinear_qkv", "dlinear_o"):
    # the flagship decode linears (csrc/gemm/decode_linear_sm100.cu) on the Llama-3-8B shapes, rotating weights (cold L2)
    from flashinfer_b200.gemm import decode_linear as dl
    H, I, D = 4096, 14336, 128
    n, k = {"dlinear_gate_up": (2 * I, H), "dlinear_down": (H, I), "dlinear_qkv": (6144, H), "dlinear_o": (H, H)}[mode]
    ws = [dl.to_block_major_k(torch.randn(n, k, device="cuda", dtype=torch.bfloat16) * 0.02) for _ in range(6)]
    a = torch.randn(64, k, device="cuda", dtype=torch.bfloat16)
    ss = torch.rand(64, device="cuda") * H + 1
    resid = torch.zeros(64, H, device="cuda", dtype=torch.bfloat16); sq = torch.zeros(64, device="cuda")
    act = torch.empty(64, I, device="cuda", dtype=torch.bfloat16)
    for w in ws:
        if mode == "dlinear_gate_up":
            dl.decode_linear(a, w, dl.EPI_GATED_SILU, out=act, row_sumsq=ss)
        elif mode in ("dlinear_down", "dlinear_o"):
            dl.decode_linear(a, w, dl.EPI_RESIDUAL, residual=resid, sumsq_out=sq)
        else:
            dl.decode_linear(a, w, out=torch.empty(64, n, device="cuda", dtype=torch.bfloat16), row_sumsq=ss)
torch.cuda.synchronize()
