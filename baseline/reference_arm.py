"""Reference arm of bench.py: the SAME Llama-3-8B decode step built ONLY from the unmodified
reference library (``baseline/_ref/flashinfer``) through its public API and stock code paths, plus
torch (cuBLAS) for the dense bf16 linears and NCCL for TP>1 — i.e. what an engine using the
reference does today.  None of flashinfer_b200's kernels, models or engine are imported here.
"""
import json
import threading
import time

import torch
import torch.distributed as dist
import torch.nn.functional as F


def run(args, rank, world, BATCH, KV_LEN, PAGE, kv_layout_fn, clock_sampler, summarise_clocks):
    import flashinfer  # resolved from baseline/_ref (bench.py put it first on sys.path)

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
    embed = rnd((vocab, hidden), gs, 1.0)
    lm_head = rnd((vshard, hidden), g, hidden ** -0.5)
    final_norm = torch.ones(hidden, device=dev, dtype=dt)
    layers = []
    for _ in range(layers_n):
        layers.append(dict(
            ln1=torch.ones(hidden, device=dev, dtype=dt), ln2=torch.ones(hidden, device=dev, dtype=dt),
            wqkv=rnd(((hq + 2 * hkv) * d, hidden), g, hidden ** -0.5), wo=rnd((hidden, hq * d), g, (hq_full * d) ** -0.5),
            wgu=rnd((2 * inter, hidden), g, hidden ** -0.5), wd=rnd((hidden, inter), g, inter_full ** -0.5),
            kc=torch.randn(n_pages, PAGE, hkv, d, device=dev, dtype=dt) * 0.5,
            vc=torch.randn(n_pages, PAGE, hkv, d, device=dev, dtype=dt) * 0.5))
    ws = torch.empty(128 << 20, dtype=torch.uint8, device=dev)
    kv_indptr, kv_indices, kv_last = indptr.to(dev), indices.to(dev), last.to(dev)
    wrapper = flashinfer.BatchDecodeWithPagedKVCacheWrapper(ws, "NHD", use_tensor_cores=True)
    wrapper.plan(kv_indptr, kv_indices, kv_last, hq, hkv, d, PAGE, q_data_type=dt, kv_data_type=dt)
    positions = torch.full((BATCH,), KV_LEN - 1, dtype=torch.int32, device=dev)
    batch_indices = torch.arange(BATCH, dtype=torch.int32, device=dev)
    tokens = torch.zeros(BATCH, dtype=torch.int64, device=dev)
    next_tokens = torch.zeros(BATCH, dtype=torch.int64, device=dev)
    res = torch.empty(BATCH, hidden, device=dev, dtype=dt)

    def reduce_add_norm(x, w):
        if world > 1:
            dist.all_reduce(x)
        flashinfer.fused_add_rmsnorm(x, res, w, eps)

    def step():
        x = embed.index_select(0, tokens)
        res.zero_()
        flashinfer.fused_add_rmsnorm(x, res, layers[0]["ln1"], eps)
        for li, l in enumerate(layers):
            qkv = F.linear(x, l["wqkv"]).view(BATCH, hq + 2 * hkv, d)
            q, k, v = qkv[:, :hq], qkv[:, hq:hq + hkv], qkv[:, hq + hkv:]
            flashinfer.apply_llama31_rope_pos_ids_inplace(q, k, positions, rope_scale=rscale, rope_theta=theta)
            flashinfer.append_paged_kv_cache(k, v, batch_indices, positions, (l["kc"], l["vc"]), kv_indices, kv_indptr,
                                             kv_last)
            attn = wrapper.run(q, (l["kc"], l["vc"]))
            x = F.linear(attn.view(BATCH, hq * d), l["wo"])
            reduce_add_norm(x, l["ln2"])
            act = flashinfer.silu_and_mul(F.linear(x, l["wgu"]))
            x = F.linear(act, l["wd"])
            reduce_add_norm(x, layers[li + 1]["ln1"] if li + 1 < layers_n else final_norm)
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
    step()  # JIT-compiles the reference modules
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
    if rank == 0:
        ms = dev_ms / args.steps
        print(json.dumps({
            "impl": "reference", "metric": "llama3_8b_paged_decode_tokens_per_s", "value": BATCH / (ms / 1e3),
            "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic (random-init weights, random KV cache, random token ids)",
            "config": {"model": "llama-3-8b", "global_batch": BATCH, "seq_len": KV_LEN, "page_size": PAGE,
                       "parallelism": f"tp{world}", "attention_backend": "fa2 use_tensor_cores=True (stock on sm100)",
                       "linear": "torch/cuBLAS", "allreduce": "NCCL" if world > 1 else None,
                       "cuda_graph": graph is not None,
                       "l2_policy": "inputs larger than L2"},
            "clocks": summarise_clocks(clk.get("rows")),
            "e2e": {"value": BATCH / (e2e_ms / args.steps / 1e3), "unit": "tokens/s", "h2d_bytes_per_step": BATCH * 8,
                    "d2h_bytes_per_step": BATCH * 8},
            "flashinfer_version": flashinfer.__version__,
        }), flush=True)
