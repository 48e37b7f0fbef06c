"""The BASELINE.json configs that are not the headline, measured through the public API of either implementation
(``--impl ours``: flashinfer_b200; ``--impl reference``: the unmodified reference from ``baseline/_ref``), same shapes,
same timing (CUDA events, cold L2 through rotating / flushed buffers, median of 7, max over ranks).  bench.py embeds the
result as the ``extra`` block of its JSON line; this file also runs standalone:

    python benchmarks/extra_configs.py --impl ours --config prefill_pod_fp8                       # BASELINE config 3 (1 GPU)
    torchrun --nproc-per-node 8 ... benchmarks/extra_configs.py --impl ours --config tp_gemm_rs   # BASELINE config 5
    torchrun --nproc-per-node 8 ... benchmarks/extra_configs.py --impl ours --config dsv3_mla_moe # BASELINE config 4

config 3  ragged causal prefill 8k + paged prefill over an fp8 (e4m3) KV cache + BatchPOD mixed batch (one 2k-token prefill
          chunk over a 8k context + 64 decode requests of 4k), GQA 32/8, head_dim 128, page 16.
config 5  Llama-3-70B row-parallel linears at B=16 x 2k = 32768 tokens: GEMM -> reduce-scatter -> +residual -> RMSNorm on the
          shard (o_proj K = 8192 / tp, down_proj K = 28672 / tp, N = 8192).
config 4  DeepSeek-V3 decode: MLA decode (128 heads, ckv 512 + kpe 64, B = 64 per GPU, kv 4096) and the expert-parallel
          NVFP4 MoE (256 experts, top-8, hidden 7168, inter 2048; 64 tokens per GPU) with dispatch / combine all-to-all.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _timed(torch, fn, world=1, dist=None, iters=7, flush=None):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = torch.tensor([sorted(ts)[len(ts) // 2]], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


# ====================================================================================================================
# config 3
# ====================================================================================================================
def prefill_pod_fp8(impl):
    import torch

    if impl == "ours":
        import flashinfer_b200 as fi
    else:
        import flashinfer as fi
    dev = "cuda"
    HQ, HKV, D, PAGE, S = 32, 8, 128, 16, 8192
    out = {"shape": {"num_qo_heads": HQ, "num_kv_heads": HKV, "head_dim": D, "page_size": PAGE, "seqlen": S, "kv_dtype": "e4m3"}}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ws = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    flops_causal = 4.0 * S * S * HQ * D * 0.5
    # ---- (a) ragged causal prefill, one 8k request, bf16
    q = torch.randn(S, HQ, D, device=dev, dtype=torch.bfloat16)
    k = torch.randn(S, HKV, D, device=dev, dtype=torch.bfloat16)
    v = torch.randn(S, HKV, D, device=dev, dtype=torch.bfloat16)
    ind = torch.tensor([0, S], dtype=torch.int32, device=dev)
    try:
        w = fi.BatchPrefillWithRaggedKVCacheWrapper(ws, "NHD")
        w.plan(ind, ind, HQ, HKV, D, causal=True, q_data_type=torch.bfloat16)
        ms = _timed(torch, lambda: w.run(q, k, v), flush=flush)
        out["ragged_prefill_8k_bf16"] = {"ms": round(ms, 4), "tflops": round(flops_causal / ms / 1e9, 1)}
    except Exception as e:  # noqa: BLE001
        out["ragged_prefill_8k_bf16"] = {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}
    # ---- (b) paged causal prefill over an fp8 KV cache
    n_pages = S // PAGE
    kc8 = (torch.randn(n_pages + 64 * 256, PAGE, HKV, D, device=dev) * 0.5).to(torch.float8_e4m3fn)
    vc8 = (torch.randn(n_pages + 64 * 256, PAGE, HKV, D, device=dev) * 0.5).to(torch.float8_e4m3fn)
    kv_indptr = torch.tensor([0, n_pages], dtype=torch.int32, device=dev)
    kv_indices = torch.arange(n_pages, dtype=torch.int32, device=dev)
    last = torch.tensor([PAGE], dtype=torch.int32, device=dev)
    try:
        wp = fi.BatchPrefillWithPagedKVCacheWrapper(ws, "NHD")
        wp.plan(ind, kv_indptr, kv_indices, last, HQ, HKV, D, PAGE, causal=True, q_data_type=torch.bfloat16,
                kv_data_type=torch.float8_e4m3fn)
        ms = _timed(torch, lambda: wp.run(q, (kc8, vc8)), flush=flush)
        out["paged_prefill_8k_fp8kv"] = {"ms": round(ms, 4), "tflops": round(flops_causal / ms / 1e9, 1)}
    except Exception as e:  # noqa: BLE001
        out["paged_prefill_8k_fp8kv"] = {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}
    # ---- (c) BatchPOD: a 2k-token prefill chunk at the end of the 8k context + 64 decode requests x 4k, fp8 KV
    CH, BD, KVD = 2048, 64, 4096
    qp = torch.randn(CH, HQ, D, device=dev, dtype=torch.bfloat16)
    qd = torch.randn(BD, HQ, D, device=dev, dtype=torch.bfloat16)
    qo_p = torch.tensor([0, CH], dtype=torch.int32, device=dev)
    ppr = KVD // PAGE
    kv_indptr_d = torch.arange(0, (BD + 1) * ppr, ppr, dtype=torch.int32, device=dev)
    kv_indices_d = (n_pages + torch.randperm(BD * ppr, device=dev)).int()
    last_d = torch.full((BD,), PAGE, dtype=torch.int32, device=dev)
    qo_d = torch.arange(0, BD + 1, dtype=torch.int32, device=dev)
    try:
        pod = fi.BatchPODWithPagedKVCacheWrapper(ws, "NHD")
        pod.plan(qo_p, kv_indptr, kv_indices, last, qo_d, kv_indptr_d, kv_indices_d, last_d, HQ, HKV, D, PAGE,
                 q_data_type=torch.bfloat16, kv_data_type=torch.float8_e4m3fn)
        ms = _timed(torch, lambda: pod.run(qp, (kc8, vc8), qd, (kc8, vc8), causal_p=True), flush=flush)
        pf = 4.0 * CH * (S - CH / 2) * HQ * D
        db = 2.0 * BD * KVD * HKV * D
        out["batch_pod_fp8kv"] = {"ms": round(ms, 4), "prefill_tflops_equiv": round(pf / ms / 1e9, 1),
                                  "decode_kv_tb_per_s_equiv": round(db / ms / 1e9, 3),
                                  "mix": f"prefill chunk {CH} over {S} ctx + {BD} decode x {KVD}"}
    except Exception as e:  # noqa: BLE001
        out["batch_pod_fp8kv"] = {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}
    return out


# ====================================================================================================================
# config 5
# ====================================================================================================================
def tp_gemm_rs(impl, rank, world):
    import torch
    import torch.distributed as dist

    hidden, inter, M = 8192, 28672, 16 * 2048
    M = M // world * world
    out = {"shape": {"tokens": M, "hidden": hidden, "tp": world}}
    if impl == "ours":
        import flashinfer_b200 as fi
        from flashinfer_b200.comm import GemmAllReduce

        comm = GemmAllReduce(None, M, hidden, torch.bfloat16)
    else:
        import flashinfer as fi
    gamma = torch.ones(hidden, device="cuda", dtype=torch.bfloat16)
    shard = torch.empty(M // world, hidden, device="cuda", dtype=torch.bfloat16)
    for name, k_full in (("o_proj", hidden), ("down_proj", inter)):
        K = k_full // world
        torch.manual_seed(rank)
        a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
        w = (torch.randn(hidden, K, device="cuda") / k_full ** 0.5).bfloat16()
        res = torch.randn(M // world, hidden, device="cuda").bfloat16()
        if impl == "ours":
            fn = lambda: comm.reduce_scatter(a, w, residual=res, rms_weight=gamma, eps=1e-5)  # noqa: E731
            path = ("tcgen05 GEMM (cta_group::2 gemm_nt) into symmetric staging, in-switch multimem.ld_reduce pull + residual + row "
                    "sum-of-squares (rs_pull_rows), rs_rmsnorm scale pass; schedule (one kernel / sequential / c pipeline chunks with "
                    "the pull on a side stream) picked per shape by the tuner; no NCCL")
        else:
            def fn():
                c = fi.mm_bf16(a, w.t()) if hasattr(fi, "mm_bf16") and os.environ.get("FIB200_REF_MM", "torch") == "flashinfer" else a @ w.t()
                dist.reduce_scatter_tensor(shard, c)
                r = res.clone()
                fi.fused_add_rmsnorm(shard, r, gamma, 1e-5)

            path = "torch/cuBLAS matmul + NCCL reduce_scatter_tensor + flashinfer.fused_add_rmsnorm"
        try:
            ms = _timed(torch, fn, world, dist)
            flops = 2.0 * M * hidden * K
            t_compute = flops / 1.64e15 * 1e3
            t_link = M * hidden * 2 * (world - 1) / world / 770e9 * 1e3
            out[name] = {"ms": round(ms, 4), "k_local": K, "roofline_ms": round(max(t_compute, t_link), 4),
                         "fraction_of_roofline": round(max(t_compute, t_link) / ms, 3), "path": path}
            if impl == "ours":  # schedule chosen by the per-shape tuner (0 one kernel | 1 GEMM then pull | c pipeline chunks) and its timings
                out[name]["schedule_ms"] = {str(k): v for k, v in next(iter(getattr(comm, "_rs_tuning_log", {}).values()), {}).items()}
                out[name]["schedule"] = next(iter(getattr(comm, "_rs_tuned", {}).values()), None)
                out[name]["nvls_multicast"] = bool(comm.use_nvls)
                comm.__dict__.pop("_rs_tuning_log", None)
                comm.__dict__.pop("_rs_tuned", None)
        except Exception as e:  # noqa: BLE001
            out[name] = {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}
        del a, w, res
    return out


# ====================================================================================================================
# config 4
# ====================================================================================================================
def dsv3_mla_moe(impl, rank, world):
    import torch
    import torch.distributed as dist

    out = {"shape": {"mla": "128 heads, ckv 512 + kpe 64, B=64/GPU, kv 4096, page 64", "moe": "256 experts top-8, hidden 7168, inter 2048, "
                     f"64 tokens/GPU, EP={world}, NVFP4 weights"}}
    dev = "cuda"
    B, KV, PAGE, H, DC, DR = 64, 4096, 64, 128, 512, 64
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # ---- MLA decode (data parallel: per-GPU number)
    try:
        if impl == "ours":
            import flashinfer_b200 as fi
        else:
            import flashinfer as fi
        ppr = KV // PAGE
        ckv = torch.randn(B * ppr, PAGE, DC, device=dev, dtype=torch.bfloat16)
        kpe = torch.randn(B * ppr, PAGE, DR, device=dev, dtype=torch.bfloat16)
        qn = torch.randn(B, H, DC, device=dev, dtype=torch.bfloat16)
        qp = torch.randn(B, H, DR, device=dev, dtype=torch.bfloat16)
        ws = torch.zeros(128 << 20, dtype=torch.uint8, device=dev)
        w = fi.mla.BatchMLAPagedAttentionWrapper(ws, backend="auto" if impl == "ours" else "fa2")
        qo = torch.arange(0, B + 1, dtype=torch.int32, device=dev)
        kvi = torch.arange(0, (B + 1) * ppr, ppr, dtype=torch.int32, device=dev)
        kvx = torch.randperm(B * ppr, device=dev).int()
        lens = torch.full((B,), KV, dtype=torch.int32, device=dev)
        w.plan(qo, kvi, kvx, lens, H, DC, DR, PAGE, False, (DC + DR) ** -0.5, torch.bfloat16, torch.bfloat16)
        ms = _timed(torch, lambda: w.run(qn, qp, ckv, kpe), world, dist if world > 1 else None, flush=flush)
        out["mla_decode"] = {"ms": round(ms, 4), "latent_tb_per_s": round(B * KV * (DC + DR) * 2 / ms / 1e9, 3)}
        del ckv, kpe
    except Exception as e:  # noqa: BLE001
        out["mla_decode"] = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
    # ---- expert-parallel NVFP4 MoE with dispatch / combine all-to-all
    E, TOPK, HID, INTER, T = 256, 8, 7168, 2048, 64
    try:
        if impl != "ours":
            # the reference's NVFP4 MoE entry points on B200: trtllm_fp4_block_scale_moe (trtllm-gen batched-GEMM cubins of the
            # reference's artifact hash) and cutlass_fused_moe / cute_dsl (CUTLASS / CuTe-DSL JIT builds of >10 min)
            from flashinfer.artifacts import ArtifactPath
            from flashinfer.jit.env import FLASHINFER_CUBIN_DIR

            need = FLASHINFER_CUBIN_DIR / ArtifactPath.TRTLLM_GEN_BMM
            if not need.exists():
                raise FileNotFoundError(f"trtllm-gen batched-GEMM cubins {ArtifactPath.TRTLLM_GEN_BMM} are not on this box (no network); "
                                        "cutlass_fused_moe needs a CUTLASS JIT build beyond the bench budget")
            raise RuntimeError("reference MoE arm not wired for this artifact set")
        from flashinfer_b200.comm import Mapping, MoeAlltoAll
        from flashinfer_b200.fused_moe.core import moe_forward_nvfp4, route
        from flashinfer_b200.quantization.fp4 import fp4_quantize

        e_local = E // world
        torch.manual_seed(5 + rank)
        x = (torch.randn(T, HID, device=dev) * 0.5).bfloat16()
        logits = torch.randn(T, E, device=dev)
        bias = torch.zeros(E, device=dev)

        def q3(n, k):  # random NVFP4 expert weights, quantised expert by expert (linear [N, K/16] block scales)
            qs, sfs = [], []
            for _ in range(e_local):
                wt = (torch.randn(n, k, device=dev) * 0.02).bfloat16()
                qq, sf = fp4_quantize(wt, torch.full((1,), 448.0 * 6.0 / 0.1, device=dev), 16, False, False)
                qs.append(qq)
                sfs.append(sf)
            return torch.stack(qs), torch.stack(sfs)

        w1q, w1s = q3(2 * INTER, HID)
        w2q, w2s = q3(HID, INTER)
        alpha = 0.1 / (448.0 * 6.0)
        ids, wts = route(logits, bias, TOPK, 2, 8, 4, 2.5, True)
        if world > 1:
            R = T
            a2a = MoeAlltoAll(Mapping(world, rank, tp_size=world, moe_ep_size=world, moe_tp_size=1), max_num_tokens=T, top_k=TOPK,
                              num_experts=E, hidden_size=HID, dtype=torch.bfloat16)

            def step():
                recv_x, recv_ids, recv_w = a2a.dispatch(ids, [x, ids, wts], R, invalid_token_expert_id=-1, expert_id_payload_index=1)
                payload = a2a.get_combine_payload_tensor_in_workspace(R, HID, torch.bfloat16).view(world * R, HID)
                moe_forward_nvfp4(recv_x.reshape(world * R, HID), recv_ids.reshape(world * R, TOPK), recv_w.reshape(world * R, TOPK),
                                  w1q, w1s, alpha, w2q, w2s, alpha, rank * e_local, E, out=payload)
                return a2a.combine(payload.view(world, R, HID), R, payload_in_workspace=True)
        else:
            def step():
                return moe_forward_nvfp4(x, ids, wts, w1q, w1s, alpha, w2q, w2s, alpha, 0, E)

        ms = _timed(torch, step, world, dist if world > 1 else None, flush=flush)
        wbytes = e_local * (2 * INTER * HID + HID * INTER) * 0.5625
        out["moe_ep_nvfp4"] = {"ms": round(ms, 4), "weight_tb_per_s_equiv": round(wbytes / ms / 1e9, 3),
                               "path": "dispatch A2A -> routing-sorted NVFP4 grouped tcgen05 GEMMs -> combine A2A" if world > 1
                               else "NVFP4 grouped tcgen05 GEMMs (single GPU, no A2A)"}
    except Exception as e:  # noqa: BLE001
        out["moe_ep_nvfp4"] = {"unavailable": f"{type(e).__name__}: {str(e)[:240]}"}
    return out


def run_extras(impl, rank, world, which=None, budget_s=600.0):
    """In-process entry used by bench.py.  Single-GPU configs run on rank 0 only when world == 1; multi-GPU configs when world > 1."""
    t0 = time.time()
    res = {}
    plan = [("prefill_pod_fp8", lambda: prefill_pod_fp8(impl), world == 1),
            ("tp_gemm_rs", lambda: tp_gemm_rs(impl, rank, world), world > 1),
            ("dsv3_mla_moe", lambda: dsv3_mla_moe(impl, rank, world), world > 1)]
    for name, fn, ok in plan:
        if not ok or (which and name not in which):
            continue
        if time.time() - t0 > budget_s:
            res[name] = {"skipped": "extras time budget exhausted"}
            continue
        try:
            res[name] = fn()
        except Exception as e:  # noqa: BLE001
            res[name] = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
        res[name]["seconds"] = round(time.time() - t0, 1)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None)
    ap.add_argument("--json-out", default=None)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    import torch

    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if args.impl == "reference":
        os.environ.setdefault("FLASHINFER_DISABLE_VERSION_CHECK", "1")
        os.environ.setdefault("FLASHINFER_WORKSPACE_BASE", os.path.join(ROOT, "baseline", "_ref_cache"))
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    else:
        sys.path.insert(0, ROOT)
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    res = run_extras(args.impl, rank, world, [args.config] if args.config else None)
    if rank == 0:
        line = json.dumps({"impl": args.impl, "n_gpus": world, "extra": res})
        print(line, flush=True)
        if args.json_out:
            with open(args.json_out, "w") as f:
                f.write(line)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
