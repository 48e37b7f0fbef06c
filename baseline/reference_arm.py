"""Reference arm of bench.py: the SAME Llama-3-8B decode step built ONLY from the unmodified reference library
(``baseline/_ref/flashinfer``) through its public API and stock code paths, plus torch (cuBLAS) for the dense bf16 linears - i.e.
what an engine using the reference does today.  None of flashinfer_b200's kernels, models or engine are imported here.

The arm uses the reference's BEST path that runs offline on this box:
  * decode attention: every backend of the reference that can be built here is tried on layer 0 - ``trtllm-gen``
    (``BatchDecodeWithPagedKVCacheWrapper(backend="trtllm-gen")``; needs cubins of the reference's own artifact hash), ``cudnn``
    (``cudnn_batch_decode_with_kv_cache``), ``fa2_tc`` (``use_tensor_cores=True``), ``fa2`` - each is checked against ``fa2_tc`` and
    timed; the fastest correct one runs the step.  What was tried, the time of each and why a backend was unavailable are
    recorded in ``config.ref_attention_candidates``.
  * TP all-reduce + residual + RMSNorm: ``trtllm_allreduce_fusion(kARResidualRMSNorm)`` (JIT-built from ``baseline/_ref``) when its
    IPC workspace can be created, else NCCL ``all_reduce`` + ``fused_add_rmsnorm``; both are timed, the faster one runs.
``--ref-attn`` / ``--ref-ar`` pin a choice.
"""
import json
import threading
import time

import torch
import torch.distributed as dist
import torch.nn.functional as F


def _time_us(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def run(args, rank, world, BATCH, KV_LEN, PAGE, kv_layout_fn, clock_sampler, summarise_clocks, extras=None):
    import flashinfer  # resolved from baseline/_ref (bench.py put it first on sys.path)

    src = getattr(args, "ref_src", "tree")
    if src == "tree":
        assert "baseline/_ref" in flashinfer.__file__.replace("\\", "/"), flashinfer.__file__
    dev = torch.device("cuda", torch.cuda.current_device())
    dt = torch.bfloat16
    hidden, inter_full, layers_n, hq_full, hkv_full, d, vocab = 4096, 14336, 32, 32, 8, 128, 128256
    eps, theta, rscale = 1e-5, 5e5, 8.0
    hq, hkv, inter = hq_full // world, hkv_full // world, inter_full // world
    vshard = (vocab + world - 1) // world
    g = torch.Generator(device=dev).manual_seed(1000 * rank)
    gs = torch.Generator(device=dev).manual_seed(0)

    def rnd(shape, gen, std):
        return (torch.randn(shape, device=dev, dtype=torch.float32, generator=gen) * std).to(dt)

    indptr, indices, last, n_pages = kv_layout_fn(BATCH, KV_LEN, PAGE, torch)
    ppr = n_pages // BATCH
    embed = rnd((vocab, hidden), gs, 1.0)
    lm_head = rnd((vshard, hidden), g, hidden ** -0.5)
    final_norm = torch.ones(hidden, device=dev, dtype=dt)
    layers = []
    for _ in range(layers_n):  # HND cache: the native layout of the trtllm-gen / cudnn backends, accepted by fa2 as well
        layers.append(dict(
            ln1=torch.ones(hidden, device=dev, dtype=dt), ln2=torch.ones(hidden, device=dev, dtype=dt),
            wqkv=rnd(((hq + 2 * hkv) * d, hidden), g, hidden ** -0.5), wo=rnd((hidden, hq * d), g, (hq_full * d) ** -0.5),
            wgu=rnd((2 * inter, hidden), g, hidden ** -0.5), wd=rnd((hidden, inter), g, inter_full ** -0.5),
            kc=torch.randn(n_pages, hkv, PAGE, d, device=dev, dtype=dt) * 0.5,
            vc=torch.randn(n_pages, hkv, PAGE, d, device=dev, dtype=dt) * 0.5))
    kv_indptr, kv_indices, kv_last = indptr.to(dev), indices.to(dev), last.to(dev)
    positions = torch.full((BATCH,), KV_LEN - 1, dtype=torch.int32, device=dev)
    batch_indices = torch.arange(BATCH, dtype=torch.int32, device=dev)
    tokens = torch.zeros(BATCH, dtype=torch.int64, device=dev)
    next_tokens = torch.zeros(BATCH, dtype=torch.int64, device=dev)
    res = torch.empty(BATCH, hidden, device=dev, dtype=dt)
    sm_scale = d ** -0.5

    # ------------------------------------------------------------------ attention backends of the reference
    def make_wrapper(backend, tc):
        if backend == "trtllm-gen":
            # the launcher would try to DOWNLOAD the cubins of the reference's own artifact hash (minutes of retries, no network):
            # check the cubin directory first and report precisely what is missing
            from flashinfer.artifacts import ArtifactPath
            from flashinfer.jit.env import FLASHINFER_CUBIN_DIR

            need = FLASHINFER_CUBIN_DIR / ArtifactPath.TRTLLM_GEN_FMHA
            if not need.exists():
                raise FileNotFoundError(f"trtllm-gen FMHA cubins {ArtifactPath.TRTLLM_GEN_FMHA} are not on this box (cubin dir "
                                        f"{FLASHINFER_CUBIN_DIR} ships other hashes; no network to fetch them)")
        ws = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)  # trtllm-gen wants a zero-initialised workspace
        w = flashinfer.BatchDecodeWithPagedKVCacheWrapper(ws, "HND", use_tensor_cores=tc, backend=backend)
        w.plan(kv_indptr, kv_indices, kv_last, hq, hkv, d, PAGE, q_data_type=dt, kv_data_type=dt)
        return lambda q, l: w.run(q, (l["kc"], l["vc"]))

    def make_cudnn():
        from flashinfer.cudnn import cudnn_batch_decode_with_kv_cache

        ws = torch.empty(128 << 20, dtype=torch.uint8, device=dev)
        block_tables = kv_indices.view(BATCH, ppr).contiguous()
        seq = torch.full((BATCH, 1, 1, 1), KV_LEN, dtype=torch.int32, device=dev)
        return lambda q, l: cudnn_batch_decode_with_kv_cache(q, l["kc"], l["vc"], sm_scale, ws, max_sequence_kv=KV_LEN,
                                                             actual_seq_lens_kv=seq, block_tables=block_tables,
                                                             is_cuda_graph_compatible=True)

    makers = {"trtllm-gen": lambda: make_wrapper("trtllm-gen", True), "cudnn": make_cudnn,
              "fa2_tc": lambda: make_wrapper("fa2", True), "fa2": lambda: make_wrapper("fa2", False)}
    want_attn = getattr(args, "ref_attn", "auto")
    # auto: fa2 on CUDA cores is left out (published 0.050 ms vs 0.022 ms for fa2_tc at B=16: never the best; one JIT build less)
    order = ["fa2_tc", "trtllm-gen", "cudnn"] if want_attn == "auto" else [want_attn]
    qprobe = torch.randn(BATCH, hq, d, device=dev, dtype=dt)
    cands, fns, oracle = {}, {}, None
    for name in order:
        try:
            fn = makers[name]()
            o = fn(qprobe, layers[0])
            torch.cuda.synchronize()
            if oracle is None:
                oracle = o.float()
            err = float((o.float() - oracle).abs().max())
            if not err < 2e-2:
                raise RuntimeError(f"output mismatch vs {order[0]}: max abs err {err:.3g}")
            us = _time_us(lambda: fn(qprobe, layers[0]))
            cands[name] = {"us_per_layer": round(us, 1)}
            fns[name] = fn
        except Exception as e:  # noqa: BLE001
            cands[name] = {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}
            torch.cuda.synchronize()
    if not fns:
        raise RuntimeError(f"no reference attention backend runs here: {cands}")
    # all ranks must agree (identical hardware: they normally do; make it exact)
    names = sorted(fns, key=lambda n: cands[n]["us_per_layer"])
    best_attn = names[0]
    if world > 1:
        pick = [best_attn]
        dist.broadcast_object_list(pick, src=0)
        best_attn = pick[0] if pick[0] in fns else best_attn
    attn = fns[best_attn]

    # ------------------------------------------------------------------ TP all-reduce (+ residual + RMSNorm)
    ar_info = {"nccl": {}}
    x_norm = torch.empty(BATCH, hidden, device=dev, dtype=dt)
    res2 = torch.empty_like(res)
    state = {"res": res, "res_alt": res2}

    def ran_nccl(x, w):
        dist.all_reduce(x)
        flashinfer.fused_add_rmsnorm(x, state["res"], w, eps)
        return x

    fusion = None
    want_ar = getattr(args, "ref_ar", "auto")
    if world > 1 and want_ar in ("auto", "trtllm_fusion"):
        try:
            import flashinfer.comm as comm

            _, ws_t = comm.trtllm_create_ipc_workspace_for_all_reduce_fusion(rank, world, BATCH, hidden, group=dist.group.WORLD)

            def ran_fusion(x, w):
                comm.trtllm_allreduce_fusion(
                    allreduce_in=x, world_size=world, world_rank=rank, token_num=BATCH, hidden_dim=hidden, workspace_ptrs=ws_t,
                    launch_with_pdl=True, trigger_completion_at_end=True, fp32_acc=False,
                    pattern_code=comm.AllReduceFusionPattern.kARResidualRMSNorm, use_oneshot=None, allreduce_out=None,
                    residual_in=state["res"], residual_out=state["res_alt"], norm_out=x_norm, quant_out=None, scale_out=None,
                    rms_gamma=w, rms_eps=eps, scale_factor=None, layout_code=None)
                state["res"], state["res_alt"] = state["res_alt"], state["res"]
                return x_norm

            # numerics vs NCCL composition, then time both
            xa = torch.randn(BATCH, hidden, device=dev, dtype=dt)
            r0 = torch.randn(BATCH, hidden, device=dev, dtype=dt)
            dist.broadcast(r0, 0)
            state["res"].copy_(r0)
            ref_x = ran_nccl(xa.clone(), final_norm).float().clone()
            state["res"].copy_(r0)
            got = ran_fusion(xa.clone(), final_norm).float()
            torch.cuda.synchronize()
            err = float((got - ref_x).abs().max())
            if not err < 5e-2:
                raise RuntimeError(f"trtllm_allreduce_fusion mismatch vs NCCL: {err:.3g}")
            fusion = ran_fusion
            xb = torch.randn(BATCH, hidden, device=dev, dtype=dt)
            ar_info["trtllm_fusion"] = {"us": round(_time_us(lambda: ran_fusion(xb, final_norm), iters=40), 1)}
            ar_info["nccl"] = {"us": round(_time_us(lambda: ran_nccl(xb, final_norm), iters=40), 1)}
        except Exception as e:  # noqa: BLE001
            ar_info["trtllm_fusion"] = {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}
            fusion = None
            torch.cuda.synchronize()
    use_fusion = False
    if fusion is not None:
        use_fusion = want_ar == "trtllm_fusion" or ar_info["trtllm_fusion"]["us"] <= ar_info["nccl"]["us"]
        flag = torch.tensor([1 if use_fusion else 0], device=dev)
        dist.broadcast(flag, 0)
        use_fusion = bool(int(flag))
    state["res"], state["res_alt"] = res, res2

    def reduce_add_norm(x, w):
        if world > 1:
            return fusion(x, w) if use_fusion else ran_nccl(x, w)
        flashinfer.fused_add_rmsnorm(x, state["res"], w, eps)
        return x

    def step():
        x = embed.index_select(0, tokens)
        state["res"].zero_()
        flashinfer.fused_add_rmsnorm(x, state["res"], layers[0]["ln1"], eps)
        for li, l in enumerate(layers):
            qkv = F.linear(x, l["wqkv"]).view(BATCH, hq + 2 * hkv, d)
            q, k, v = qkv[:, :hq], qkv[:, hq:hq + hkv], qkv[:, hq + hkv:]
            flashinfer.apply_llama31_rope_pos_ids_inplace(q, k, positions, rope_scale=rscale, rope_theta=theta)
            flashinfer.append_paged_kv_cache(k, v, batch_indices, positions, (l["kc"], l["vc"]), kv_indices, kv_indptr,
                                             kv_last, kv_layout="HND")
            a = attn(q.contiguous() if best_attn in ("trtllm-gen", "cudnn") else q, l)
            x = F.linear(a.view(BATCH, hq * d), l["wo"])
            x = reduce_add_norm(x, l["ln2"])
            act = flashinfer.silu_and_mul(F.linear(x, l["wgu"]))
            x = F.linear(act, l["wd"])
            x = reduce_add_norm(x, layers[li + 1]["ln1"] if li + 1 < layers_n else final_norm)
        logits = F.linear(x, lm_head)
        if world == 1:
            next_tokens.copy_(torch.argmax(logits, dim=-1))
        else:
            val, idx = torch.max(logits.float(), dim=-1)
            vals = [torch.empty_like(val) for _ in range(world)]
            idxs = [torch.empty_like(idx) for _ in range(world)]
            dist.all_gather(vals, val)
            dist.all_gather(idxs, idx + rank * vshard)
            best = torch.stack(vals, 0).argmax(0)
            next_tokens.copy_(torch.stack(idxs, 0).gather(0, best[None])[0])
        return next_tokens

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    tok_pin = torch.randint(0, vocab, (BATCH,), dtype=torch.int64).pin_memory()
    out_pin = torch.empty(BATCH, dtype=torch.int64).pin_memory()
    tokens.copy_(tok_pin)
    step()  # JIT-compiles the remaining reference modules
    torch.cuda.synchronize()
    if getattr(args, "eager_steps", 0) > 0:
        for _ in range(args.eager_steps):
            step()
        torch.cuda.synchronize()
        return
    graph = None
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            step()
            step()  # an even number of steps keeps the fusion path's residual ping-pong aligned with the captured graph
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
    except Exception:  # noqa: BLE001  (NCCL capture etc.) -> eager
        graph = None
        torch.cuda.synchronize()

    def replay():
        if graph is not None:
            graph.replay()
        else:
            step()

    for _ in range(args.warmup):
        replay()
    barrier()
    stop_evt, clk = threading.Event(), {}
    th = threading.Thread(target=clock_sampler, args=(stop_evt, clk), daemon=True)
    th.start()
    time.sleep(0.3)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s_, e_ in evs:
        s_.record()
        replay()
        e_.record()
    barrier()
    dev_ms = sum(s_.elapsed_time(e_) for s_, e_ in evs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tokens.copy_(tok_pin, non_blocking=True)
        replay()
        out_pin.copy_(next_tokens, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        tok_pin.copy_(out_pin)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    stop_evt.set()
    if clk.get("proc") is not None:
        try:
            clk["proc"].kill()
        except Exception:  # noqa: BLE001
            pass
    tm = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = tm.tolist()
    extra = None
    if extras is not None:
        del layers, embed, lm_head  # the extra configs allocate their own tensors
        torch.cuda.empty_cache()
        extra = extras()
    if rank == 0:
        ms = dev_ms / args.steps
        kv_bytes = BATCH * KV_LEN * hkv_full * d * 2 * 2 * layers_n
        print(json.dumps({
            "impl": "reference", "metric": "llama3_8b_paged_decode_tokens_per_s", "value": BATCH / (ms / 1e3),
            "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic (random-init weights, random KV cache, random token ids)",
            "config": {"model": "llama-3-8b", "global_batch": BATCH, "seq_len": KV_LEN, "page_size": PAGE,
                       "parallelism": f"tp{world}", "kv_layout": "HND", "cuda_graph": graph is not None,
                       "l2_policy": "inputs larger than L2 (34 GB KV + 16 GB weights streamed per step)",
                       "attention_kv_tb_per_s_equiv": kv_bytes / world / (ms / 1e3) / 1e12,
                       "attention_backend": best_attn, "ref_attention_candidates": cands, "linear": "torch/cuBLAS",
                       "allreduce": ("trtllm_allreduce_fusion(kARResidualRMSNorm)" if use_fusion else "NCCL + fused_add_rmsnorm")
                       if world > 1 else None, "ref_allreduce_candidates": ar_info if world > 1 else None},
            "clocks": summarise_clocks(clk.get("rows")),
            "e2e": {"value": BATCH / (e2e_ms / args.steps / 1e3), "unit": "tokens/s", "h2d_bytes_per_step": BATCH * 8,
                    "d2h_bytes_per_step": BATCH * 8},
            "flashinfer_version": flashinfer.__version__,
            "extra": extra,
            "reference_source": "baseline/_ref (pip install of /root/reference)" if src == "tree" else f"installed wheel ({flashinfer.__file__})",
        }), flush=True)
