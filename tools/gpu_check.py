#!/usr/bin/env python
"""Developer GPU check runner: every case runs in its own subprocess under a timeout so that a hung
kernel (mbarrier deadlock) cannot take the whole gpurun call down.  Results -> gpurun_out/check.json.

    python tools/gpu_check.py                 # run all cases
    python tools/gpu_check.py --case gemm     # run one case in-process
"""
import argparse
import json
import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _time_ms(fn, iters=20, warmup=5, flush=None):
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def case_gemm():
    import torch
    import flashinfer_b200 as fi
    from flashinfer_b200.gemm import linear

    torch.manual_seed(0)
    res = {}
    shapes = [(128, 128, 64), (128, 256, 128), (256, 512, 512), (1024, 4096, 4096), (300, 1000 // 8 * 8, 520),
              (1, 4096, 4096), (8, 512, 1024), (16, 6144, 4096), (64, 4096, 14336), (64, 28672, 4096), (100, 384, 2048),
              (4096, 4096, 4096), (8192, 8192, 8192)]
    for (m, n, k) in shapes:
        for dt in (torch.bfloat16,):
            x = torch.randn(m, k, device="cuda", dtype=dt)
            w = torch.randn(n, k, device="cuda", dtype=dt) / (k ** 0.5)
            y = linear(x, w)
            ref = (x.float() @ w.float().t())
            err = (y.float() - ref).abs().max().item()
            tol = 2e-2 * ref.abs().max().item() + 1e-2
            ok = err <= tol
            res[f"{m}x{n}x{k}"] = {"err": err, "tol": tol, "ok": bool(ok)}
            print(f"gemm {m}x{n}x{k} err={err:.4g} tol={tol:.4g} {'OK' if ok else 'FAIL'}", flush=True)
    # bias + fp16
    x = torch.randn(77, 512, device="cuda", dtype=torch.float16)
    w = torch.randn(264, 512, device="cuda", dtype=torch.float16) / 22
    b = torch.randn(264, device="cuda", dtype=torch.float16)
    y = linear(x, w, b)
    ref = x.float() @ w.float().t() + b.float()
    err = (y.float() - ref).abs().max().item()
    res["fp16_bias"] = {"err": err, "ok": bool(err < 3e-2)}
    print("gemm fp16+bias err", err, flush=True)
    # perf
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for (m, n, k) in [(8192, 8192, 8192), (4096, 4096, 4096), (64, 28672, 4096), (64, 4096, 14336), (64, 6144, 4096), (64, 4096, 4096)]:
        x = torch.randn(m, k, device="cuda", dtype=torch.bfloat16)
        w = torch.randn(n, k, device="cuda", dtype=torch.bfloat16)
        ms = _time_ms(lambda: linear(x, w), flush=flush)
        ms_t = _time_ms(lambda: torch.nn.functional.linear(x, w), flush=flush)
        tf = 2 * m * n * k / ms / 1e9
        gbs = (m * k + n * k + m * n) * 2 / ms / 1e6
        res[f"perf_{m}x{n}x{k}"] = {"ms": ms, "tflops": tf, "gbs": gbs, "cublas_ms": ms_t}
        print(f"gemm perf {m}x{n}x{k}: {ms:.4f} ms {tf:.1f} TFLOP/s {gbs:.0f} GB/s  (cuBLAS {ms_t:.4f} ms)", flush=True)
    return res


def _make_paged(batch, kv_lens, hkv, d, ps, layout, dtype, device="cuda"):
    import torch

    npages = [(l + ps - 1) // ps for l in kv_lens]
    total = sum(npages)
    indptr = torch.tensor([0] + list(torch.tensor(npages).cumsum(0).tolist()), dtype=torch.int32)
    indices = torch.randperm(total + 3)[:total].int()
    last = torch.tensor([(l - 1) % ps + 1 if l > 0 else 0 for l in kv_lens], dtype=torch.int32)
    shape = (total + 3, ps, hkv, d) if layout == "NHD" else (total + 3, hkv, ps, d)
    kc = torch.randn(shape, device=device, dtype=dtype)
    vc = torch.randn(shape, device=device, dtype=dtype)
    return indptr, indices, last, kc, vc


def case_decode():
    import torch
    import flashinfer_b200 as fi
    from flashinfer_b200 import reference

    torch.manual_seed(0)
    res = {}
    ws = torch.empty(128 << 20, dtype=torch.uint8, device="cuda")
    cfgs = [
        # (kv_lens, hq, hkv, ps, layout, dtype)
        ([128], 4, 1, 16, "NHD", torch.bfloat16),
        ([37, 128, 300], 8, 2, 16, "NHD", torch.bfloat16),
        ([1, 17, 1000, 4096], 32, 8, 16, "NHD", torch.bfloat16),
        ([513, 64], 32, 8, 16, "HND", torch.float16),
        ([700, 90, 5], 16, 16, 32, "NHD", torch.bfloat16),
        ([700, 90, 5], 32, 4, 8, "HND", torch.bfloat16),
        ([333], 8, 8, 128, "NHD", torch.float16),
        ([1000, 33], 8, 1, 256, "NHD", torch.bfloat16),
        ([77, 200], 8, 2, 1, "NHD", torch.bfloat16),
        ([4096] * 16, 32, 8, 16, "NHD", torch.bfloat16),
    ]
    for (kv_lens, hq, hkv, ps, layout, dt) in cfgs:
        name = f"kv{kv_lens[:3]}x{len(kv_lens)}_h{hq}/{hkv}_ps{ps}_{layout}_{str(dt)[6:]}"
        try:
            B = len(kv_lens)
            indptr, indices, last, kc, vc = _make_paged(B, kv_lens, hkv, 128, ps, layout, dt)
            q = torch.randn(B, hq, 128, device="cuda", dtype=dt)
            w = fi.BatchDecodeWithPagedKVCacheWrapper(ws, layout)
            w.plan(indptr, indices, last, hq, hkv, 128, ps, q_data_type=dt)
            o, lse = w.run(q, (kc, vc), return_lse=True)
            torch.cuda.synchronize()
            qo = torch.arange(B + 1, dtype=torch.int32)
            o_ref, lse_ref = reference.batch_paged_attention_ref(
                q, qo, kc, vc, indptr, indices.cuda(), last, layout, True)
            err = (o.float() - o_ref.float()).abs().max().item()
            lerr = (lse - lse_ref).abs().max().item()
            ok = err < 2e-2 and lerr < 2e-2
            res[name] = {"err": err, "lse_err": lerr, "ok": bool(ok), "counts": w._plan_counts.tolist()}
            print(f"decode {name}: err={err:.4g} lse_err={lerr:.4g} {'OK' if ok else 'FAIL'}", flush=True)
        except Exception as e:  # noqa: BLE001
            res[name] = {"ok": False, "exc": repr(e)}
            print(f"decode {name}: EXC {e!r}", flush=True)
            traceback.print_exc()
    return res


def case_decode_perf():
    import torch
    import flashinfer_b200 as fi

    torch.manual_seed(0)
    res = {}
    ws = torch.empty(128 << 20, dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for (B, kv, hq, hkv, ps) in [(64, 4096, 32, 8, 16), (16, 1024, 64, 8, 16), (64, 1024, 32, 8, 16), (256, 4096, 32, 8, 16), (1, 8192, 32, 8, 16), (64, 16384, 32, 8, 16)]:
        kv_lens = [kv] * B
        indptr, indices, last, kc, vc = _make_paged(B, kv_lens, hkv, 128, ps, "NHD", torch.bfloat16)
        q = torch.randn(B, hq, 128, device="cuda", dtype=torch.bfloat16)
        w = fi.BatchDecodeWithPagedKVCacheWrapper(ws, "NHD")
        w.plan(indptr, indices, last, hq, hkv, 128, ps, q_data_type=torch.bfloat16)
        out = torch.empty_like(q)
        ms = _time_ms(lambda: w.run(q, (kc, vc), out=out), flush=flush if B * kv * hkv * 512 < (200 << 20) else None)
        byts = B * kv * hkv * 128 * 2 * 2 + 2 * q.numel() * 2
        flops = 4 * B * hq * kv * 128
        res[f"B{B}_kv{kv}_h{hq}/{hkv}"] = {"ms": ms, "tbs": byts / ms / 1e9, "tflops": flops / ms / 1e9}
        print(f"decode perf B={B} kv={kv} h={hq}/{hkv}: {ms:.4f} ms  {byts / ms / 1e9:.3f} TB/s  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)
    return res


def case_prefill():
    import torch
    import flashinfer_b200 as fi
    from flashinfer_b200 import reference

    torch.manual_seed(0)
    res = {}
    ws = torch.empty(128 << 20, dtype=torch.uint8, device="cuda")
    # (q_lens, kv_lens, hq, hkv, causal, dtype, paged page_size or 0)
    cfgs = [
        ([128], [128], 1, 1, False, torch.bfloat16, 0),
        ([256], [256], 2, 1, True, torch.bfloat16, 0),
        ([100, 300, 17], [100, 300, 17], 8, 2, True, torch.bfloat16, 0),
        ([33, 257], [500, 1000], 4, 4, True, torch.float16, 0),
        ([512], [2048], 8, 2, False, torch.bfloat16, 0),
        ([100, 300, 17], [150, 300, 400], 8, 2, True, torch.bfloat16, 16),
        ([64, 129], [1000, 129], 4, 1, True, torch.float16, 32),
        ([1000], [1000], 32, 8, True, torch.bfloat16, 0),
    ]
    for (q_lens, kv_lens, hq, hkv, causal, dt, ps) in cfgs:
        name = f"q{q_lens}_kv{kv_lens}_h{hq}/{hkv}_c{int(causal)}_ps{ps}_{str(dt)[6:]}"
        try:
            B = len(q_lens)
            qo = torch.tensor([0] + torch.tensor(q_lens).cumsum(0).tolist(), dtype=torch.int32)
            q = torch.randn(sum(q_lens), hq, 128, device="cuda", dtype=dt)
            if ps == 0:
                kvi = torch.tensor([0] + torch.tensor(kv_lens).cumsum(0).tolist(), dtype=torch.int32)
                k = torch.randn(sum(kv_lens), hkv, 128, device="cuda", dtype=dt)
                v = torch.randn(sum(kv_lens), hkv, 128, device="cuda", dtype=dt)
                w = fi.BatchPrefillWithRaggedKVCacheWrapper(ws)
                w.plan(qo, kvi, hq, hkv, 128, causal=causal, q_data_type=dt)
                o, lse = w.run(q, k, v, return_lse=True)
                torch.cuda.synchronize()
                outs, lses = [], []
                for b in range(B):
                    o_r, l_r = reference.attention_ref(q[qo[b]:qo[b + 1]], k[kvi[b]:kvi[b + 1]], v[kvi[b]:kvi[b + 1]], causal)
                    outs.append(o_r); lses.append(l_r)
                o_ref, lse_ref = torch.cat(outs), torch.cat(lses)
            else:
                indptr, indices, last, kc, vc = _make_paged(B, kv_lens, hkv, 128, ps, "NHD", dt)
                w = fi.BatchPrefillWithPagedKVCacheWrapper(ws)
                w.plan(qo, indptr, indices, last, hq, hkv, 128, ps, causal=causal, q_data_type=dt)
                o, lse = w.run(q, (kc, vc), return_lse=True)
                torch.cuda.synchronize()
                o_ref, lse_ref = reference.batch_paged_attention_ref(q, qo, kc, vc, indptr, indices.cuda(), last, "NHD", causal)
            err = (o.float() - o_ref.float()).abs().max().item()
            lerr = (lse - lse_ref).abs().max().item()
            ok = err < 3e-2 and lerr < 2e-2
            res[name] = {"err": err, "lse_err": lerr, "ok": bool(ok)}
            print(f"prefill {name}: err={err:.4g} lse_err={lerr:.4g} {'OK' if ok else 'FAIL'}", flush=True)
        except Exception as e:  # noqa: BLE001
            res[name] = {"ok": False, "exc": repr(e)}
            print(f"prefill {name}: EXC {e!r}", flush=True)
            traceback.print_exc()
    return res


def case_prefill_perf():
    import torch
    import flashinfer_b200 as fi

    res = {}
    ws = torch.empty(128 << 20, dtype=torch.uint8, device="cuda")
    for (B, s, hq, hkv, causal) in [(1, 8192, 32, 8, True), (4, 8192, 32, 8, True), (16, 1024, 32, 8, True), (2, 16384, 32, 8, True), (8, 4096, 32, 8, False)]:
        qo = torch.arange(0, (B + 1) * s, s, dtype=torch.int32)
        q = torch.randn(B * s, hq, 128, device="cuda", dtype=torch.bfloat16)
        k = torch.randn(B * s, hkv, 128, device="cuda", dtype=torch.bfloat16)
        v = torch.randn(B * s, hkv, 128, device="cuda", dtype=torch.bfloat16)
        w = fi.BatchPrefillWithRaggedKVCacheWrapper(ws)
        w.plan(qo, qo, hq, hkv, 128, causal=causal, q_data_type=torch.bfloat16)
        out = torch.empty_like(q)
        ms = _time_ms(lambda: w.run(q, k, v, out=out), iters=10, warmup=3)
        flops = 4 * B * hq * s * s * 128 * (0.5 if causal else 1.0)
        res[f"B{B}_s{s}_c{int(causal)}"] = {"ms": ms, "tflops": flops / ms / 1e9}
        print(f"prefill perf B={B} s={s} causal={causal}: {ms:.3f} ms {flops / ms / 1e9:.1f} TFLOP/s", flush=True)
    return res


def _graph_time_us(fn, reps=20, iters=5):
    """Per-call GPU time of `fn` measured with a CUDA graph holding `reps` calls (no python overhead)."""
    import torch

    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g.replay()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) * 1e3 / reps)
    return best


def case_gemm_small():
    """Decode-shape GEMMs (M=64) timed inside CUDA graphs; weights rotate over > L2 worth of copies."""
    import torch
    from flashinfer_b200.gemm import linear

    res = {}
    for (m, n, k) in [(64, 6144, 4096), (64, 4096, 4096), (64, 28672, 4096), (64, 4096, 14336), (64, 128256, 4096),
                      (16, 4096, 4096), (128, 8192, 8192)]:
        ncopies = max(2, int(300e6 // (n * k * 2)) + 1)
        ws = [torch.randn(n, k, device="cuda", dtype=torch.bfloat16) for _ in range(ncopies)]
        x = torch.randn(m, k, device="cuda", dtype=torch.bfloat16)
        out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        state = {"i": 0}

        def ours():
            linear(x, ws[state["i"] % ncopies], out=out)
            state["i"] += 1

        def cublas():
            torch.nn.functional.linear(x, ws[state["i"] % ncopies], out=None)
            state["i"] += 1

        ref = x.float() @ ws[0].float().t()
        state["i"] = 0
        y = linear(x, ws[0])
        err = (y.float() - ref).abs().max().item() / ref.abs().max().item()
        t_o = _graph_time_us(ours)
        t_c = _graph_time_us(cublas)
        res[f"{m}x{n}x{k}"] = {"ours_us": t_o, "cublas_us": t_c, "rel_err": err, "gbs": n * k * 2 / t_o / 1e3}
        print(f"gemm_small {m}x{n}x{k}: ours {t_o:.2f} us ({n * k * 2 / t_o / 1e3:.0f} GB/s)  cuBLAS {t_c:.2f} us  rel_err {err:.3g}", flush=True)
    return res


def case_mla_perf():
    import torch
    from flashinfer_b200.mla import BatchMLAPagedAttentionWrapper

    res = {}
    for (B, kv, H, ps) in [(16, 1024, 128, 32), (64, 4096, 128, 64), (128, 8192, 128, 64), (1, 8192, 128, 64)]:
        npg = (kv + ps - 1) // ps
        kvp = torch.arange(0, (B + 1) * npg, npg, dtype=torch.int32)
        idx = torch.randperm(B * npg).int()
        ckv = torch.randn(B * npg, ps, 512, device="cuda", dtype=torch.bfloat16)
        kpe = torch.randn(B * npg, ps, 64, device="cuda", dtype=torch.bfloat16)
        qn = torch.randn(B, H, 512, device="cuda", dtype=torch.bfloat16)
        qp = torch.randn(B, H, 64, device="cuda", dtype=torch.bfloat16)
        w = BatchMLAPagedAttentionWrapper(torch.empty(256 << 20, dtype=torch.uint8, device="cuda"))
        w.plan(torch.arange(B + 1, dtype=torch.int32), kvp, idx, torch.full((B,), kv, dtype=torch.int32), H, 512, 64, ps,
               True, 0.07, torch.bfloat16, torch.bfloat16)
        out = torch.empty(B, H, 512, device="cuda", dtype=torch.bfloat16)
        us = _graph_time_us(lambda: w.run(qn, qp, ckv, kpe, out=out), reps=5)
        byts = B * kv * 576 * 2
        flops = 2 * B * H * kv * (576 + 512)
        res[f"B{B}_kv{kv}"] = {"us": us, "tbs": byts / us / 1e6, "tflops": flops / us / 1e6}
        print(f"mla perf B={B} kv={kv} H={H}: {us:.1f} us  {byts / us / 1e6:.3f} TB/s  {flops / us / 1e6:.1f} TFLOP/s", flush=True)
    return res


def case_moe_perf():
    """MoE pipeline + grouped GEMM throughput (DeepSeek-V3-like and Mixtral-like shapes) vs a torch loop."""
    import torch
    from flashinfer_b200.fused_moe import moe_forward, route
    from flashinfer_b200.gemm import grouped_gemm_tiles

    res = {}
    for name, (T, E, K, H, I) in {"dsv3_ep8": (4096, 32, 8, 7168, 2048), "mixtral": (4096, 8, 2, 4096, 14336),
                                  "dsv3_decode": (128, 32, 8, 7168, 2048)}.items():
        x = (torch.randn(T, H, device="cuda") * 0.5).bfloat16()
        w1 = (torch.randn(E, 2 * I, H, device="cuda") / H ** 0.5).bfloat16()
        w2 = (torch.randn(E, H, I, device="cuda") / I ** 0.5).bfloat16()
        ids, w = route(torch.randn(T, E, device="cuda"), None, K, 1)
        ms = _time_ms(lambda: moe_forward(x, ids, w, w1, w2), iters=10, warmup=3)
        flops = 2.0 * T * K * H * I * 3
        # dense lower bound: one cuBLAS GEMM pair with the same FLOPs
        a = torch.randn(T * K, H, device="cuda").bfloat16()
        ms_cublas = _time_ms(lambda: ((a @ w1[0].t())[:, :I].contiguous() @ w2[0].t()), iters=10, warmup=3)
        res[name] = {"ms": ms, "tflops": flops / ms / 1e9, "cublas_dense_same_flops_ms": ms_cublas}
        # grouped GEMM alone (FC1 shape), all tiles full
        rows = (T * K + 127) // 128 * 128
        ap = torch.randn(rows, H, device="cuda").bfloat16()
        te = (torch.arange(rows // 128, device="cuda") % E).int()
        msg = _time_ms(lambda: grouped_gemm_tiles(ap, w1, te, None), iters=10, warmup=3)
        res[name]["fc1_grouped_ms"] = msg
        res[name]["fc1_grouped_tflops"] = 2.0 * rows * H * 2 * I / msg / 1e9
        del x, w1, w2, a, ap
    return res


def case_lowp_perf():
    """fp8 / mxfp8 / nvfp4 GEMM throughput (8192^3 and a decode shape) with forced and heuristic N tiles."""
    import torch
    import flashinfer_b200 as fi

    res = {}
    for (m, n, k) in [(8192, 8192, 8192), (4096, 14336, 4096), (64, 14336, 4096)]:
        a = torch.randn(m, k, device="cuda", dtype=torch.bfloat16)
        w = torch.randn(n, k, device="cuda", dtype=torch.bfloat16)
        key = f"{m}x{n}x{k}"
        res[key] = {}
        flops = 2.0 * m * n * k
        a8, w8 = a.to(torch.float8_e4m3fn), w.to(torch.float8_e4m3fn)
        one = torch.ones(1, device="cuda")
        aq8, asf8 = fi.mxfp8_quantize(a)
        wq8, wsf8 = fi.mxfp8_quantize(w)
        g = torch.tensor(1.0, device="cuda")
        aq4, asf4 = fi.nvfp4_quantize(a, g)
        wq4, wsf4 = fi.nvfp4_quantize(w, g)
        for bn in [0, 128, 192, 224, 256]:
            os.environ["FIB200_LOWP_BN"] = str(bn)
            try:
                t8 = _time_ms(lambda: fi.mm_fp8(a8, w8.t(), one), iters=10, warmup=3)
                tm8 = _time_ms(lambda: fi.mm_mxfp8(aq8, wq8.t(), asf8, wsf8), iters=10, warmup=3)
                t4 = _time_ms(lambda: fi.mm_fp4(aq4, wq4.t(), asf4, wsf4, one), iters=10, warmup=3)
                res[key][f"bn{bn}"] = {"fp8_tflops": flops / t8 / 1e9, "mxfp8_tflops": flops / tm8 / 1e9,
                                       "nvfp4_tflops": flops / t4 / 1e9, "fp8_ms": t8, "mxfp8_ms": tm8, "nvfp4_ms": t4}
            except Exception as e:  # noqa: BLE001
                res[key][f"bn{bn}"] = repr(e)[:300]
        os.environ.pop("FIB200_LOWP_BN", None)
        tb = _time_ms(lambda: torch.matmul(a, w.t()), iters=10, warmup=3)
        res[key]["cublas_bf16_tflops"] = flops / tb / 1e9
        try:
            ts = _time_ms(lambda: torch._scaled_mm(a8, w8.t(), scale_a=one, scale_b=one, out_dtype=torch.bfloat16), iters=10, warmup=3)
            res[key]["cublaslt_fp8_tflops"] = flops / ts / 1e9
        except Exception as e:  # noqa: BLE001
            res[key]["cublaslt_fp8_tflops"] = repr(e)[:200]
    return res


CASES = {"lowp_perf": case_lowp_perf, "moe_perf": case_moe_perf, "mla_perf": case_mla_perf, "gemm_small": case_gemm_small, "prefill": case_prefill, "prefill_perf": case_prefill_perf, "gemm": case_gemm, "decode": case_decode, "decode_perf": case_decode_perf}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--cases", default=None, help="comma list for the driver mode")
    ap.add_argument("--timeout", type=int, default=240)
    args = ap.parse_args()
    if args.case:
        out = CASES[args.case]()
        print("RESULT_JSON " + json.dumps(out))
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    names = args.cases.split(",") if args.cases else list(CASES)
    summary = {}
    for name in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", name], capture_output=True,
                               text=True, timeout=args.timeout, cwd=ROOT)
            tail = (p.stdout + "\n" + p.stderr)[-6000:]
            js = None
            for line in p.stdout.splitlines():
                if line.startswith("RESULT_JSON "):
                    js = json.loads(line[len("RESULT_JSON "):])
            summary[name] = {"rc": p.returncode, "sec": time.time() - t0, "result": js, "tail": tail if js is None else tail[-1500:]}
            print(f"=== {name}: rc={p.returncode} ({time.time() - t0:.1f}s)\n{p.stdout[-4000:]}\n{p.stderr[-2000:]}", flush=True)
        except subprocess.TimeoutExpired as e:
            summary[name] = {"rc": "timeout", "sec": time.time() - t0,
                             "tail": ((e.stdout or b"").decode(errors="replace") if isinstance(e.stdout, bytes) else (e.stdout or ""))[-4000:]}
            print(f"=== {name}: TIMEOUT\n{summary[name]['tail']}", flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "check.json"), "w") as f:
        json.dump(summary, f, indent=1)


if __name__ == "__main__":
    main()
