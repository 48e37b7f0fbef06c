#!/usr/bin/env python
"""Flagship benchmark: Llama-3-8B batched decode step over a paged KV cache (BASELINE.json config #2
shapes: 32/8 heads, head_dim 128, page_size 16, batch 64, kv_len 4096, bf16) with tensor parallelism
over N GPUs (TP all-reduce + residual-add + RMSNorm fusion, BASELINE.json config #5 pattern).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nproc-per-node 8 ... bench.py --gpus 8 --steps 20 --warmup 5
    python bench.py --impl reference ...          # unmodified reference (baseline/_ref) arm

One step = embedding -> 32 x (RMSNorm, QKV GEMM, RoPE, paged-KV append, paged decode attention,
O GEMM, [TP all-reduce]+add+RMSNorm, gate/up GEMM, SiLU*mul, down GEMM, [TP all-reduce]+add+RMSNorm)
-> LM head GEMM -> greedy sampling, for 64 requests each holding 4096 cached tokens.
Metric: generated tokens/s aggregated over the job ("strong" scaling: total work is fixed).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))

BATCH, KV_LEN, PAGE = 64, 4096, 16


def _clock_sampler(stop_evt, out):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:  # noqa: BLE001
        return
    out["proc"] = p
    dev = os.environ.get("LOCAL_RANK", "0")
    for line in p.stdout:
        parts = [x.strip() for x in line.split(",")]
        if len(parts) >= 9 and parts[0] == dev:
            out.setdefault("rows", []).append(parts)
        if stop_evt.is_set():
            break
    try:
        p.kill()
    except Exception:  # noqa: BLE001
        pass


def _summarise_clocks(rows):
    if not rows:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = sorted(int(float(r[1])) for r in rows)
    reasons = set()
    for r in rows:
        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
            if v.lower().startswith("active"):
                reasons.add(name)
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(rows[0][2])), "reasons": sorted(reasons),
            "samples": len(rows)}


def _kv_layout(batch, kv_len, page, torch):
    ppr = (kv_len + page - 1) // page
    indptr = torch.arange(0, (batch + 1) * ppr, ppr, dtype=torch.int32)
    g = torch.Generator().manual_seed(7)
    indices = torch.randperm(batch * ppr, generator=g).int()
    last = torch.full((batch,), (kv_len - 1) % page + 1, dtype=torch.int32)
    return indptr, indices, last, batch * ppr


def run_ours(args, rank, world):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import flashinfer_b200 as fi
    from flashinfer_b200 import jit
    from flashinfer_b200.models.llama import LlamaConfig, LlamaDecodeEngine

    cfg = LlamaConfig.llama3_8b()
    comm = None
    if world > 1:
        from flashinfer_b200.comm import TPCommunicator
        comm = TPCommunicator(dist.group.WORLD, max_tokens=BATCH, hidden=cfg.hidden_size, dtype=torch.bfloat16)
    indptr, indices, last, n_pages = _kv_layout(BATCH, KV_LEN, PAGE, torch)
    eng = LlamaDecodeEngine(cfg, BATCH, n_pages, PAGE, tp_rank=rank, tp_size=world, comm=comm, fused=not args.unfused,
                            kv_layout=args.kv_layout)
    eng.fill_kv_random()
    eng.plan(indptr, indices, last)
    # host-side inputs in pinned memory (e2e path) + device staging
    tok_pin = torch.randint(0, cfg.vocab_size, (BATCH,), dtype=torch.int64).pin_memory()
    out_pin = torch.empty(BATCH, dtype=torch.int64).pin_memory()
    eng.tokens.copy_(tok_pin)
    c0 = jit.native_launch_count()
    eng.step()
    torch.cuda.synchronize()
    native_per_step = jit.native_launch_count() - c0
    # torch-launched kernels per step: op-by-op path index_select, zero_ (+ argmax on one GPU)
    torch_per_step = 0 if eng.fused else (3 if world == 1 else 2)  # fused path: every launch is native (argmax included)
    if args.eager_steps > 0:  # ncu / profiler mode: plain eager launches, no timing contract
        for _ in range(args.eager_steps):
            eng.step()
        torch.cuda.synchronize()
        return
    eng.capture(warmup=1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > L2 (126 MB); KV itself is 34 GB/step
    for _ in range(args.warmup):
        eng.replay()
    barrier()
    stop_evt, clk = threading.Event(), {}
    th = threading.Thread(target=_clock_sampler, args=(stop_evt, clk), daemon=True)
    th.start()
    time.sleep(0.3)
    # ---- device-timed steps (CUDA events, graph replay, inputs > L2 every step) ----
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s, e in evs:
        s.record()
        eng.replay()
        e.record()
    barrier()
    dev_ms = sum(s.elapsed_time(e) for s, e in evs)
    # ---- end-to-end steps through the public engine API: H2D inputs, step, D2H result ----
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.tokens.copy_(tok_pin, non_blocking=True)
        eng.replay()
        out_pin.copy_(eng.next_tokens, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        tok_pin.copy_(out_pin)  # autoregressive feed-back on the host
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    stop_evt.set()
    if clk.get("proc") is not None:
        try:
            clk["proc"].kill()
        except Exception:  # noqa: BLE001
            pass
    tm = torch.tensor([dev_ms, e2e_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = tm.tolist()
    del flush
    res = None
    if rank == 0:
        ms_per_step = dev_ms / args.steps
        value = BATCH / (ms_per_step / 1e3)
        e2e = BATCH / (e2e_ms / args.steps / 1e3)
        kv_bytes = BATCH * KV_LEN * cfg.num_kv_heads * cfg.head_dim * 2 * 2 * cfg.num_layers
        res = {
            "metric": "llama3_8b_paged_decode_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic (random-init weights, random KV cache, random token ids)", "impl": "flashinfer_b200",
            # same keys as the reference arm's config (the driver compares the two dicts)
            "config": {"model": "llama-3-8b", "global_batch": BATCH, "seq_len": KV_LEN, "page_size": PAGE,
                       "parallelism": f"tp{world}", "kv_layout": args.kv_layout, "cuda_graph": True,
                       "l2_policy": "inputs larger than L2 (34 GB KV + 16 GB weights streamed per step)",
                       "attention_kv_tb_per_s_equiv": kv_bytes / world / (ms_per_step / 1e3) / 1e12,
                       "attention_backend": "flashinfer_b200 tcgen05 paged decode (decode_sm100)",
                       "ref_attention_candidates": None,
                       "linear": ("flashinfer_b200 decode_linear_sm100 (RMSNorm / RoPE+append / SwiGLU / residual epilogues)"
                                  if eng.fused else "flashinfer_b200 gemm_sm100"),
                       "allreduce": ((("two-shot Lamport all-reduce inside the O / down GEMM epilogue (reduce-scatter push to the row owner, "
                                       "all-gather push of the new residual rows, sentinel polling; "
                                       + ("multimem.st" if eng.tp_fused.mc_recv else "peer stores: no NVLS multicast mapping on this box") + ")"
                                       if eng.tp_fused.algo == 2 else
                                       "one-shot push all-reduce inside the O / down GEMM epilogue (every rank's strip into slot [rank] of all "
                                       "ranks, sentinel polling; "
                                       + ("multimem.st" if eng.tp_fused.mc_recv else "peer stores: no NVLS multicast mapping on this box") + ")")
                                      if eng.fused else "in-kernel all-reduce + add + RMSNorm kernel") if world > 1 else None),
                       "ref_allreduce_candidates": None},
            "clocks": _summarise_clocks(clk.get("rows")),
            "e2e": {"value": e2e, "unit": "tokens/s", "h2d_bytes_per_step": BATCH * 8, "d2h_bytes_per_step": BATCH * 8},
            "gpu_launches": int((native_per_step + torch_per_step) * args.steps),
            "gpu_launches_native_per_step": int(native_per_step),
            # the other BASELINE.json configs, measured after the timed region (benchmarks/extra_configs.py; same code both arms)
            "extra": None,
        }
    _emit_guarded(res, lambda: _run_extras(args, "ours", rank, world))


def _emit_guarded(res, extras_fn, limit_s=None):
    """The headline line must never be lost to the extras: they run under a wall-clock guard.  If they raise, the error goes into the
    ``extra`` block; if they hang (a collective that never completes at some world size), every rank gives up after ``limit_s``
    seconds, rank 0 prints the headline with an error note, and the processes exit cleanly."""
    limit_s = float(os.environ.get("FIB200_BENCH_EXTRAS_LIMIT", "300")) if limit_s is None else limit_s
    lock, state = threading.Lock(), {"printed": False}

    def emit(extra):
        with lock:
            if state["printed"]:
                return
            state["printed"] = True
            if res is not None:
                res["extra"] = extra
                print(json.dumps(res), flush=True)

    def bail():
        emit({"error": f"extras exceeded {limit_s:.0f} s and were abandoned"})
        sys.stdout.flush()
        os._exit(0)

    timer = threading.Timer(limit_s, bail)
    timer.daemon = True
    timer.start()
    try:
        extra = extras_fn()
    except BaseException as e:  # noqa: BLE001
        extra = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
    timer.cancel()
    emit(extra)


def _run_extras(args, impl, rank, world):
    """BASELINE.json configs 3 / 4 / 5 (benchmarks/extra_configs.py) after the headline measurement; never fatal.  The reference's
    single-GPU config needs several JIT builds: it runs in a subprocess with a time limit so that the headline line is never at risk."""
    if args.no_extras or os.environ.get("FIB200_BENCH_EXTRAS", "1") == "0":
        return None
    try:
        import torch

        torch.cuda.empty_cache()
        if impl == "reference" and world == 1:
            import tempfile

            if args.ref_src != "tree":
                return None
            out = os.path.join(tempfile.gettempdir(), f"fib200_ref_extras_{os.getpid()}.json")
            cmd = [sys.executable, os.path.join(ROOT, "benchmarks", "extra_configs.py"), "--impl", "reference", "--config",
                   "prefill_pod_fp8", "--json-out", out]
            try:
                subprocess.run(cmd, timeout=float(os.environ.get("FIB200_REF_EXTRAS_TIMEOUT", "540")), stdout=subprocess.DEVNULL,
                               stderr=subprocess.DEVNULL, check=False)
                with open(out) as f:
                    return json.load(f)["extra"]
            except subprocess.TimeoutExpired:
                return {"prefill_pod_fp8": {"unavailable": "reference JIT builds + run exceeded the extras time limit"}}
        sys.path.insert(0, ROOT)
        from benchmarks.extra_configs import run_extras

        return run_extras(impl, rank, world)
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {str(e)[:200]}"}


def run_reference(args, rank, world):
    ref = os.path.join(ROOT, "baseline", "_ref")
    if args.ref_src == "tree" and not os.path.isdir(os.path.join(ref, "flashinfer")):
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref not installed"}))
        return
    os.environ.setdefault("FLASHINFER_DISABLE_VERSION_CHECK", "1")
    os.environ.setdefault("FLASHINFER_WORKSPACE_BASE", os.path.join(ROOT, "baseline", "_ref_cache" if args.ref_src == "tree" else "_wheel_cache"))
    if args.ref_src == "tree":
        sys.path.insert(0, ref)
    try:
        from baseline.reference_arm import run as ref_run
        ref_run(args, rank, world, BATCH, KV_LEN, PAGE, _kv_layout, _clock_sampler, _summarise_clocks,
                extras=lambda: _run_extras(args, "reference", rank, world))
    except Exception as e:  # noqa: BLE001
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": f"{type(e).__name__}: {str(e)[:300]}"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--eager-steps", type=int, default=0, help="run N eager steps and exit (for ncu)")
    ap.add_argument("--ref-attn", default="auto", choices=["auto", "fa2", "fa2_tc", "trtllm-gen", "cudnn"],
                    help="reference arm: decode attention backend (auto = fastest that runs)")
    ap.add_argument("--ref-ar", default="auto", choices=["auto", "nccl", "trtllm_fusion"],
                    help="reference arm: TP all-reduce path (auto = faster of the two)")
    ap.add_argument("--ref-src", default="tree", choices=["tree", "wheel"],
                    help="reference arm: tree = baseline/_ref (the unmodified /root/reference, default and the driver's arm); wheel = "
                         "the pip-installed flashinfer-python 0.6.11.post2 + flashinfer-cubin (diagnostic: the only build whose "
                         "trtllm-gen cubins exist offline)")
    ap.add_argument("--kv-layout", default=os.environ.get("FIB200_BENCH_KV_LAYOUT", "NHD"), choices=["NHD", "HND"],
                    help="ours: paged KV-cache layout (HND: one contiguous 4 KB chunk per (page, head))")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra BASELINE configs (extra block of the JSON line)")
    ap.add_argument("--unfused", action="store_true", help="ours: op-by-op decode path (round-1 composition) instead of decode_linear")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch

    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sys.path.insert(0, ROOT)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world)
    if world > 1:
        import torch.distributed as dist
        try:  # (after a failed extra the CUDA context may be unusable: the JSON line is out, leave quietly)
            dist.barrier()
            dist.destroy_process_group()
        except Exception:  # noqa: BLE001
            pass


if __name__ == "__main__":
    main()
