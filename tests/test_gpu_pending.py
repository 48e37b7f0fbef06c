"""GPU tests written AFTER this round's GPU budget was spent: they have never run on a B200.  They carry the ``gpu_pending`` marker
(not ``gpu``), so the round-end ``pytest -m gpu`` does not pick up unverified assertions; the first GPU call of the next round should
be ``python -m pytest tests/test_gpu_pending.py -m gpu_pending -q`` - whatever passes is re-marked ``gpu`` and moved next to its
module's tests."""
import math
import os

import pytest
import torch

import flashinfer_b200 as fi

pytestmark = pytest.mark.gpu_pending


def _ragged(lens, hq, hkv, d=128, dv=None, seed=0):
    torch.manual_seed(seed)
    ind = torch.tensor([0] + lens).cumsum(0).int()
    n = int(ind[-1])
    mk = lambda h, dd: torch.randn(n, h, dd, device="cuda", dtype=torch.bfloat16)  # noqa: E731
    return ind, mk(hq, d), mk(hkv, d), mk(hkv, dv or d)


def _attn_ref(q, k, v, ind, sm, bias_fn=None, sinks=None):
    out = torch.zeros(q.shape[0], q.shape[1], v.shape[-1], device=q.device)
    g = q.shape[1] // k.shape[1]
    for i in range(ind.numel() - 1):
        s, e = int(ind[i]), int(ind[i + 1])
        n = e - s
        lg = torch.einsum("qhd,khd->hqk", q[s:e].float(), k[s:e].float().repeat_interleave(g, 1)) * sm
        pos = torch.arange(n, device=q.device)
        if bias_fn is not None:
            lg = lg + bias_fn(pos[None, :] - pos[:, None])
        lg = lg.masked_fill(pos[None, :] > pos[:, None], float("-inf"))
        if sinks is not None:
            p = torch.softmax(torch.cat([lg, sinks.float()[:, None, None].expand(-1, n, 1)], -1), -1)[..., :-1]
        else:
            p = torch.softmax(lg, -1)
        out[s:e] = torch.einsum("hqk,khd->qhd", p, v[s:e].float().repeat_interleave(g, 1))
    return out


def test_rpe_variant_on_the_tcgen05_prefill_kernel():
    from flashinfer_b200.cute_dsl.attention import BatchPrefillCuteDSLWrapper, RPEAttention

    ind, q, k, v = _ragged([200, 77, 513], 8, 2)
    table = torch.randn(8, 2 * 64 + 1, device="cuda") * 0.5
    w = BatchPrefillCuteDSLWrapper(torch.empty(64 << 20, dtype=torch.uint8, device="cuda"))
    w.plan(ind, ind, 8, 2, 128, causal=True, sm_scale=128 ** -0.5, q_data_type=torch.bfloat16, kv_data_type=torch.bfloat16,
           variant=RPEAttention(table, 64))
    out = w.run(q, k, v)
    ref = _attn_ref(q, k, v, ind, 128 ** -0.5, bias_fn=lambda rel: table[:, (rel + 64).clamp(0, 128)])
    torch.testing.assert_close(out.float(), ref, atol=3e-2, rtol=3e-2)


def test_cute_dsl_prefill_wrapper_builtin_variants():
    from flashinfer_b200.cute_dsl.attention import ALiBiAttention, AttentionWithSink, BatchPrefillCuteDSLWrapper

    ind, q, k, v = _ragged([130, 300], 8, 4, seed=1)
    w = BatchPrefillCuteDSLWrapper(torch.empty(64 << 20, dtype=torch.uint8, device="cuda"))
    slopes = torch.rand(8, device="cuda") * 0.2
    w.plan(ind, ind, 8, 4, 128, causal=True, sm_scale=128 ** -0.5, q_data_type=torch.bfloat16, kv_data_type=torch.bfloat16,
           variant=ALiBiAttention(slopes))
    torch.testing.assert_close(w.run(q, k, v).float(), _attn_ref(q, k, v, ind, 128 ** -0.5, bias_fn=lambda rel: slopes[:, None, None] * rel),
                               atol=3e-2, rtol=3e-2)
    sinks = torch.randn(8, device="cuda")
    w.plan(ind, ind, 8, 4, 128, causal=True, sm_scale=128 ** -0.5, q_data_type=torch.bfloat16, kv_data_type=torch.bfloat16,
           variant=AttentionWithSink(sinks))
    torch.testing.assert_close(w.run(q, k, v).float(), _attn_ref(q, k, v, ind, 128 ** -0.5, sinks=sinks), atol=3e-2, rtol=3e-2)


def test_ragged_deepseek_entry_point_applies_attention_sinks():
    ind, q, k, v = _ragged([90, 260], 16, 16, d=192, dv=128, seed=2)
    sinks = torch.randn(16, device="cuda")
    sm = 192 ** -0.5
    lens = (ind[1:] - ind[:-1]).cuda()
    out = fi.prefill.trtllm_ragged_attention_deepseek(q, k, v, torch.empty(64 << 20, dtype=torch.uint8, device="cuda"), lens, 260, 260, sm, 1.0, -1.0,
                                                      2, -1, ind.cuda(), ind.cuda(), attention_sinks=sinks)
    torch.testing.assert_close(out.float(), _attn_ref(q, k, v, ind, sm, sinks=sinks), atol=3e-2, rtol=3e-2)


def test_cute_dsl_mla_wrapper_and_xqa_batch_entry_points():
    from flashinfer_b200.cute_dsl.attention import BatchMLADecodeCuteDSLWrapper
    from flashinfer_b200.trace import templates as T

    for tpl, api in ((T.trtllm_batch_decode_mla_trace, fi.mla.trtllm_batch_decode_with_kv_cache_mla),
                     (T.xqa_batch_decode_mla_trace, fi.mla.xqa_batch_decode_with_kv_cache_mla),
                     (T.xqa_batch_decode_trace, fi.decode.xqa_batch_decode_with_kv_cache)):
        kw = tpl.make_inputs(device="cuda", seed=3, batch_size=5)
        ref = tpl.run_reference({k_: (v_.clone() if isinstance(v_, torch.Tensor) else v_) for k_, v_ in kw.items()})
        torch.testing.assert_close(api(**kw).float(), ref.float(), atol=3e-2, rtol=3e-2)
    kw = T.trtllm_batch_decode_mla_trace.make_inputs(device="cuda", seed=4, batch_size=6)
    w = BatchMLADecodeCuteDSLWrapper(torch.zeros(64 << 20, dtype=torch.int8, device="cuda"))
    w.plan(512, 64, kw["query"].shape[2], kw["kv_cache"].shape[2], torch.bfloat16)
    got = w.run(kw["query"], kw["kv_cache"], kw["block_tables"], kw["seq_lens"], kw["max_seq_len"], kw["bmm1_scale"])
    torch.testing.assert_close(got.float(), T.trtllm_batch_decode_mla_trace.run_reference(kw).float(), atol=3e-2, rtol=3e-2)


def test_every_trace_template_against_the_native_kernels():
    """tests/test_trace_templates.py on the device: every template's reference against the API it is bound to, with CUDA inputs
    (the template suite had only ever run through the CPU paths).  Failures are collected, not raised one by one."""
    import importlib
    import sys

    sys.path.insert(0, os.path.dirname(__file__))
    os.environ["FIB200_TRACE_TEST_DEVICE"] = "cuda"
    mod = importlib.import_module("test_trace_templates")
    mod = importlib.reload(mod)
    failures = []
    for (m, p, tpl), tid in zip(mod.BINDINGS, mod._IDS):
        if tpl.init is None:
            continue
        try:
            mod.test_reference_matches_api(m, p, tpl)
            torch.cuda.synchronize()
        except Exception as exc:  # noqa: BLE001
            failures.append(f"{tid}: {type(exc).__name__}: {str(exc)[:200]}")
    os.environ["FIB200_TRACE_TEST_DEVICE"] = "cpu"
    assert not failures, "\n".join(failures)


def test_cupti_timing_excludes_launch_gaps():
    from flashinfer_b200.testing import bench_gpu_time_with_cuda_event, bench_gpu_time_with_cupti

    a = torch.randn(2048, 2048, device="cuda", dtype=torch.bfloat16)

    def fn():
        for _ in range(4):
            torch.mm(a, a)

    dev = bench_gpu_time_with_cupti(fn, dry_run_iters=3, repeat_iters=20)
    wall = bench_gpu_time_with_cuda_event(fn, dry_run_iters=3, repeat_iters=20)
    med = lambda xs: sorted(xs)[len(xs) // 2]  # noqa: E731
    assert len(dev) == 20 and 0 < med(dev) <= med(wall) * 1.05


def test_moe_building_blocks_on_cuda():
    """moe_utils on the device: static-shape sort, native activation + finalize kernels, composed against the fused bf16 MoE."""
    from flashinfer_b200.fused_moe import moe_utils as mu
    from flashinfer_b200.fused_moe.core import moe_forward

    torch.manual_seed(0)
    t, k, e, h, inter, tile = 257, 4, 16, 512, 256, 128
    x = (torch.randn(t, h, device="cuda") * 0.5).to(torch.bfloat16)
    w1 = (torch.randn(e, 2 * inter, h, device="cuda") / h ** 0.5).to(torch.bfloat16)
    w2 = (torch.randn(e, h, inter, device="cuda") / inter ** 0.5).to(torch.bfloat16)
    scales, ids = torch.topk(torch.softmax(torch.randn(t, e, device="cuda"), -1), k)
    ids = ids.int()
    te, lim, e2p, p2e, total, ntiles = mu.moe_sort(ids, scales, e, k, tile_tokens_dim=tile)
    rows = mu.get_max_num_permuted_tokens(t, k, e, tile)
    xp = torch.empty(rows, h, dtype=torch.bfloat16, device="cuda")
    mu.moe_permute(x, xp, lim, p2e, ntiles, rows, k, tile)
    h1 = torch.zeros(rows, 2 * inter, dtype=torch.bfloat16, device="cuda")
    for tl in range(int(ntiles)):
        h1[tl * tile:(tl + 1) * tile] = xp[tl * tile:(tl + 1) * tile] @ w1[int(te[tl])].t()
    a = torch.empty(rows, inter, dtype=torch.bfloat16, device="cuda")
    mu.moe_swiglu(h1, a, lim, ntiles, rows, tile)
    h2 = torch.zeros(rows, h, dtype=torch.bfloat16, device="cuda")
    for tl in range(int(ntiles)):
        h2[tl * tile:(tl + 1) * tile] = a[tl * tile:(tl + 1) * tile] @ w2[int(te[tl])].t()
    out = torch.empty(t, h, dtype=torch.bfloat16, device="cuda")
    mu.moe_unpermute(h2, out, e2p, scales, t, k)
    ref = moe_forward(x, ids, scales.float(), w1, w2)
    torch.testing.assert_close(out.float(), ref.float(), atol=5e-2, rtol=5e-2)


def _pa_worker(rank, world, port, errs):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        from flashinfer_b200.parallel_attention import ParallelAttention, get_parallel_groups

        g = torch.Generator().manual_seed(0)
        H, S, D = 8, 4096, 128
        q, k, v = (torch.randn(H, S, D, generator=g).to(torch.bfloat16).cuda() for _ in range(3))
        ref = torch.nn.functional.scaled_dot_product_attention(q[None].float(), k[None].float(), v[None].float())[0]
        shard = lambda t_: t_.chunk(world, dim=1)[rank].contiguous()  # noqa: E731
        for mode in ("ulysses", "ring"):
            rg, ug = get_parallel_groups(world if mode == "ulysses" else 1, world if mode == "ring" else 1)
            out = ParallelAttention("sm100", ug, rg, fuse_qkv=True).run(shard(q), shard(k), shard(v), "HND")
            errs[(rank, mode)] = float((out.float() - shard(ref)).abs().max())
    finally:
        dist.destroy_process_group()


def test_parallel_attention_nccl_two_gpus():
    import socket

    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    errs = mp.Manager().dict()
    mp.spawn(_pa_worker, args=(2, port, errs), nprocs=2, join=True)
    assert len(errs) == 4 and max(errs.values()) < 3e-2, dict(errs)


@pytest.mark.parametrize("family", ["deepseek", "mixtral_8x7b", "qwen3_30b_a3b", "gemma2_9b"])
def test_model_engines_on_cuda_match_their_cpu_runs(family):
    """models.deepseek / models.transformer: the same random-init engine on CUDA (native kernels) and on CPU (eager paths)."""
    from flashinfer_b200.models import DeepSeekConfig, DeepSeekDecodeEngine, TransformerConfig, TransformerDecodeEngine

    page_size = 64 if family == "deepseek" else 16
    lens = [70, 1, 200, 33]
    per = [(n + page_size - 1) // page_size for n in lens]
    g = torch.Generator().manual_seed(0)
    ids = torch.randperm(sum(per) + 2, generator=g)[: sum(per)].int()
    indptr = torch.tensor([0] + list(torch.tensor(per).cumsum(0)), dtype=torch.int32)
    last = torch.tensor([(n - 1) % page_size + 1 for n in lens], dtype=torch.int32)
    tokens = torch.randint(0, 300, (4,), generator=g)
    logits = {}
    for dev in ("cpu", "cuda"):
        if family == "deepseek":
            eng = DeepSeekDecodeEngine(DeepSeekConfig.tiny(), 4, sum(per) + 2, page_size, dev, torch.bfloat16, seed=1)
            caches = ("ckv_cache", "kpe_cache")
        else:
            eng = TransformerDecodeEngine(getattr(TransformerConfig, family)().tiny(), 4, sum(per) + 2, page_size, dev, torch.bfloat16, seed=1)
            caches = ("k_cache", "v_cache")
        fill = torch.Generator().manual_seed(2)
        for l in eng.layers:
            for name in caches:
                l[name].copy_((torch.randn(l[name].shape, generator=fill) * 0.5).to(torch.bfloat16))
        eng.plan(indptr, ids, last)
        eng.tokens.copy_(tokens)
        eng.step()
        logits[dev] = eng.logits.float().cpu()
    cos = torch.nn.functional.cosine_similarity(logits["cpu"].flatten(), logits["cuda"].flatten(), dim=0)
    assert cos > 0.995, float(cos)


@pytest.mark.parametrize("family", ["mamba2", "gdn"])
def test_recurrent_engines_on_cuda_match_their_cpu_runs(family):
    from flashinfer_b200.models import GDNConfig, GDNDecodeEngine, Mamba2Config, Mamba2DecodeEngine

    slots = torch.tensor([3, 0, 5], dtype=torch.int32)
    g = torch.Generator().manual_seed(0)
    toks = [torch.randint(0, 200, (3,), generator=g) for _ in range(4)]
    logits = {}
    for dev in ("cpu", "cuda"):
        eng = (Mamba2DecodeEngine(Mamba2Config.tiny(), 6, dev, torch.bfloat16, seed=1) if family == "mamba2"
               else GDNDecodeEngine(GDNConfig.tiny(), 6, dev, torch.bfloat16, seed=1))
        eng.plan(slots)
        for t in toks:
            eng.tokens.copy_(t)
            eng.step()
        logits[dev] = eng.logits.float().cpu()
    assert torch.nn.functional.cosine_similarity(logits["cpu"].flatten(), logits["cuda"].flatten(), dim=0) > 0.995


def test_rope_llama_mode_on_cuda():
    """pos_encoding_mode="ROPE_LLAMA" (rotate, then the plain kernel) on the device against rotating q / k by hand."""
    from flashinfer_b200.attention.rope_on_the_fly import rotate_rows

    torch.manual_seed(0)
    q, k, v = (torch.randn(n, h, 128, device="cuda", dtype=torch.bfloat16) for n, h in ((100, 8), (300, 2), (300, 2)))
    got = fi.single_prefill_with_kv_cache(q, k, v, causal=True, pos_encoding_mode="ROPE_LLAMA")
    want = fi.single_prefill_with_kv_cache(rotate_rows(q, torch.arange(200, 300), 1.0, 1e4), rotate_rows(k, torch.arange(300), 1.0, 1e4), v, causal=True)
    torch.testing.assert_close(got.float(), want.float(), atol=2e-2, rtol=2e-2)
    got = fi.single_decode_with_kv_cache(q[0], k, v, pos_encoding_mode="ROPE_LLAMA")
    want = fi.single_decode_with_kv_cache(rotate_rows(q[:1], torch.tensor([299]), 1.0, 1e4)[0], rotate_rows(k, torch.arange(300), 1.0, 1e4), v)
    torch.testing.assert_close(got.float(), want.float(), atol=2e-2, rtol=2e-2)


def test_gdn_reference_state_conventions_on_cuda():
    """K-last pools, intermediate states, pool-form pretranspose decode and prefill checkpoints on the device against the CPU run
    of the same calls (the CUDA kernel works on K-major copies of the touched slots)."""
    from flashinfer_b200.gdn import chunk_gated_delta_rule, gated_delta_rule_decode_pretranspose, gated_delta_rule_mtp

    torch.manual_seed(0)
    B, T, H, HV, K, V = 3, 4, 4, 8, 128, 128
    mk = lambda *s: torch.randn(*s)  # noqa: E731
    x = dict(q=mk(B, T, H, K).bfloat16(), k=mk(B, T, H, K).bfloat16(), v=mk(B, T, HV, V).bfloat16(), a=mk(B, T, HV).bfloat16(), b=mk(B, T, HV).bfloat16())
    A_log, dt_bias, pool = mk(HV) * 0.5, mk(HV) * 0.1, mk(6, HV, V, K) * 0.1
    idx = torch.tensor([4, 0, 5], dtype=torch.int32)
    res = {}
    for dev in ("cpu", "cuda"):
        d = {n: t.to(dev) for n, t in x.items()}
        p, buf = pool.clone().to(dev), torch.zeros(B, T, HV, V, K, device=dev)
        o, _ = gated_delta_rule_mtp(d["q"], d["k"], d["v"], p, idx.to(dev), A_log.to(dev), d["a"], dt_bias.to(dev), d["b"],
                                    intermediate_states_buffer=buf, disable_state_update=False)
        p2 = pool.clone().to(dev)
        o1, _ = gated_delta_rule_decode_pretranspose(d["q"][:, :1], d["k"][:, :1], d["v"][:, :1], None, A_log.to(dev), d["a"][:, :1], dt_bias.to(dev), d["b"][:, :1],
                                                     initial_state=p2, initial_state_indices=torch.tensor([4, -1, 5], device=dev))
        res[dev] = [t.float().cpu() for t in (o, p, buf, o1, p2)]
    for got, want in zip(res["cuda"], res["cpu"]):
        torch.testing.assert_close(got, want, atol=3e-2, rtol=3e-2)
    lens = [150, 64, 130]
    cu = torch.tensor([0] + torch.tensor(lens).cumsum(0).tolist(), dtype=torch.int32)
    total = int(cu[-1])
    q, k, v = mk(total, H, K).bfloat16(), mk(total, H, K).bfloat16(), mk(total, HV, V).bfloat16()
    g, beta, init = torch.exp(-torch.rand(total, HV) * 0.3), torch.rand(total, HV), mk(3, HV, V, K) * 0.1
    starts = torch.tensor([0, 2, 3, 5])
    res = {}
    for dev in ("cpu", "cuda"):
        ck = torch.zeros(5, HV, V, K, device=dev)
        o, s = chunk_gated_delta_rule(q.to(dev), k.to(dev), v.to(dev), g.to(dev), beta.to(dev), initial_state=init.to(dev), output_final_state=True,
                                      cu_seqlens=cu.to(dev), use_qk_l2norm_in_kernel=True, state_checkpoints=ck, checkpoint_cu_starts=starts,
                                      checkpoint_every_n_tokens=64)
        res[dev] = [t.float().cpu() for t in (o, s, ck)]
    for got, want in zip(res["cuda"], res["cpu"]):
        torch.testing.assert_close(got, want, atol=3e-2, rtol=3e-2)


def test_mamba_spec_decoding_forms_on_cuda():
    """Intermediate-state caching and the varlen / num_accepted_tokens form of selective_state_update on the device against the
    CPU run of the same call (rounds of single-token launches of the CUDA kernel on an fp32 working copy)."""
    from flashinfer_b200.mamba import selective_state_update

    torch.manual_seed(0)
    H, DIM, DS, G, POOL, T = 8, 64, 128, 2, 16, 4
    A, D, dtb = -torch.rand(H, DIM, DS), torch.randn(H, DIM), torch.randn(H, DIM) * 0.1
    state = (torch.randn(POOL, H, DIM, DS) * 0.1).bfloat16()
    mk = lambda *s: torch.randn(*s).bfloat16()  # noqa: E731
    x, dt, B, C, z = mk(3, T, H, DIM), mk(3, T, H, DIM), mk(3, T, G, DS), mk(3, T, G, DS), mk(3, T, H, DIM)
    idx = torch.tensor([7, -1, 2], dtype=torch.int32)
    lens = [1, 4, 2]
    cu = torch.tensor([0, 1, 5, 7], dtype=torch.int32)
    src = torch.arange(12, dtype=torch.int32).view(3, 4)
    acc = torch.tensor([1, 3, 2])
    res = {}
    for dev in ("cpu", "cuda"):
        mv = lambda t: t.to(dev)  # noqa: E731
        st, buf = mv(state.clone()), torch.zeros(3, T, H, DIM, DS, device=dev)
        y = selective_state_update(st, mv(x), mv(dt), mv(A), mv(B), mv(C), mv(D), mv(z), mv(dtb), True, state_batch_indices=mv(idx),
                                   intermediate_states_buffer=buf, cache_steps=T)
        st2 = mv(state.clone())
        f = lambda t: mv(t.reshape(12, *t.shape[2:])[:7])  # noqa: E731
        y2 = selective_state_update(st2, f(x), f(dt), mv(A), f(B), f(C), mv(D), f(z), mv(dtb), True, state_batch_indices=mv(src),
                                    cu_seqlens=mv(cu), num_accepted_tokens=mv(acc), cache_steps=4)
        res[dev] = [t.float().cpu() for t in (y, st, buf, y2, st2)]
    for got, want in zip(res["cuda"], res["cpu"]):
        torch.testing.assert_close(got, want, atol=3e-2, rtol=3e-2)


def _mc_worker(rank, world, port, errs):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        from flashinfer_b200.comm.mixed_comm import MixedCommHandler, MixedCommMode, MixedCommOp, run_mixed_comm

        x = lambda r, rows: (torch.arange(rows * 1024, dtype=torch.float32).view(rows, 1024) % 7 * (r + 1)).bfloat16().cuda()  # noqa: E731
        for tp, dp, ops in ((world, None, (MixedCommOp.ALLREDUCE,)), (None, world, (MixedCommOp.ALLGATHER, MixedCommOp.REDUCESCATTER))):
            h = MixedCommHandler(rank, world, rank, world, 0, 1, tp, dp, None, None, torch.bfloat16, torch.device("cuda", rank), max_tokens=256, hidden=1024)
            assert MixedCommMode.FUSED_OPT_WAITS_MC in h.valid_mode_list
            for op in ops:
                rows = 64 * world if op == MixedCommOp.REDUCESCATTER else 64
                fused = run_mixed_comm(op, h, x(rank, rows))                                   # autotune -> the NVLS kernels
                nccl = run_mixed_comm(op, h, x(rank, rows), None, MixedCommMode.NCCL_ONE)
                errs[(rank, op.name)] = float((fused.float() - nccl.float()).abs().max())
            h.shutdown()
    finally:
        dist.destroy_process_group()


def test_mixed_comm_fused_modes_two_gpus():
    import socket

    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    errs = mp.Manager().dict()
    mp.spawn(_mc_worker, args=(2, port, errs), nprocs=2, join=True)
    assert len(errs) == 6 and max(errs.values()) < 0.5, dict(errs)
