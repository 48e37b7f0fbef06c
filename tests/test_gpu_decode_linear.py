"""decode_linear_sm100 kernels (tcgen05 small-M GEMM with fused epilogues) vs the fp32 PyTorch oracle of the same op
(flashinfer_b200/gemm/decode_linear.py CPU path).  Reference strategy: tests/gemm/test_tgv_gemm.py (small-M GEMM vs torch),
tests/attention/test_rope.py (rope + append vs composed ops), tests/comm/test_trtllm_allreduce_fusion.py (fused AR + residual +
norm vs NCCL + torch)."""
import socket

import pytest
import torch

from flashinfer_b200.gemm import decode_linear as dl
from flashinfer_b200.gemm.dense import interleave_gate_up

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp(min=1e-6))


def _xw(m, n, k, dtype, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = (torch.randn(m, k, device="cuda", generator=g) * 0.5).to(dtype)
    w = (torch.randn(n, k, device="cuda", generator=g) / k ** 0.5).to(dtype)
    return x, w


@pytest.mark.parametrize("m,n,k,bn,s", [(64, 4096, 4096, 0, 0), (64, 6144, 4096, 0, 0), (17, 4096, 2048, 64, 2), (64, 4096, 512, 0, 0),
                                        (1, 1024, 256, 32, 1), (64, 4096, 14336, 0, 0), (33, 4112, 1024, 48, 1), (64, 768, 4096, 0, 0),
                                        (64, 2048, 1024, 128, 2), (64, 512, 8192, 256, 2), (64, 4096, 4096, 128, 4), (50, 4096, 14336, 128, 4),
                                        (64, 6144, 4096, 192, 4), (20, 1024, 512, 64, 4)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_plain_and_rowscale(m, n, k, bn, s, dtype):
    x, w = _xw(m, n, k, dtype)
    ref = x.float() @ w.float().t()
    out = dl.decode_linear(x, w, dl.EPI_PLAIN, bn=bn, split_k=s)
    assert _rel(out, ref) < 1.5e-2
    bias = torch.randn(n, device="cuda").to(dtype)
    ss = torch.rand(64, device="cuda") * k + 1.0
    out = dl.decode_linear(x, w, dl.EPI_PLAIN, bias=bias, row_sumsq=ss, norm_dim=k, eps=1e-5, bn=bn, split_k=s)
    ref2 = ref * torch.rsqrt(ss[:m] / k + 1e-5)[:, None] + bias.float()
    assert _rel(out, ref2) < 1.5e-2
    # BlockMajorK weight layout ([K / 64, N, 64], one contiguous chunk per TMA box)
    out = dl.decode_linear(x, dl.to_block_major_k(w), dl.EPI_PLAIN, bn=bn, split_k=s)
    assert _rel(out, ref) < 1.5e-2


@pytest.mark.parametrize("m,n,k,bn,s", [(64, 28672, 4096, 0, 0), (40, 3584, 4096, 0, 0), (64, 2048, 2048, 64, 2), (5, 256, 512, 32, 1)])
def test_gated_silu(m, n, k, bn, s):
    x, w = _xw(m, n, k, torch.bfloat16, 1)
    wi = interleave_gate_up(w)
    ss = torch.rand(64, device="cuda") * k + 1.0
    out = dl.decode_linear(x, wi, dl.EPI_GATED_SILU, row_sumsq=ss, norm_dim=k, bn=bn, split_k=s)
    h = (x.float() @ w.float().t()) * torch.rsqrt(ss[:m] / k + 1e-5)[:, None]
    ref = torch.nn.functional.silu(h[:, : n // 2]) * h[:, n // 2:]
    assert out.shape == (m, n // 2)
    assert _rel(out, ref) < 2e-2


@pytest.mark.parametrize("m,n,k,bn,s", [(64, 4096, 4096, 0, 0), (64, 4096, 14336, 0, 0), (9, 1024, 512, 32, 1), (64, 4096, 1792, 0, 0)])
def test_residual_sumsq(m, n, k, bn, s):
    x, w = _xw(m, n, k, torch.bfloat16, 2)
    res = torch.randn(m, n, device="cuda").bfloat16()
    res0 = res.clone()
    ssq = torch.full((64,), 3.0, device="cuda")
    dl.decode_linear(x, w, dl.EPI_RESIDUAL, residual=res, sumsq_out=ssq, bn=bn, split_k=s)
    ref = (res0.float() + x.float() @ w.float().t())
    assert _rel(res, ref) < 1.5e-2
    torch.testing.assert_close(ssq[:m], 3.0 + res.float().pow(2).sum(-1), rtol=2e-3, atol=1e-2)
    assert float((ssq[m:] - 3.0).abs().max()) == 0.0 if m < 64 else True


@pytest.mark.parametrize("interleave", [False, True])
@pytest.mark.parametrize("m,hq,hkv,k", [(64, 32, 8, 4096), (7, 4, 1, 512), (64, 4, 1, 4096)])
def test_rope_append(m, hq, hkv, k, interleave):
    d, page, n_pages = 128, 16, 40
    n = (hq + 2 * hkv) * d
    x, w = _xw(m, n, k, torch.bfloat16, 3)
    wp = w if interleave else dl.permute_rope_rows(w, hq, hkv, d)
    pos = torch.randint(0, 3000, (m,), device="cuda", dtype=torch.int32)
    cs = torch.zeros(64, d, device="cuda")
    inv = torch.pow(torch.tensor(5e5), -torch.arange(0, d // 2, device="cuda").float() * 2 / d)
    ang = pos.float()[:, None] * inv[None]
    cs[:m, : d // 2], cs[:m, d // 2:] = torch.cos(ang), torch.sin(ang)
    kc = torch.zeros(n_pages, page, hkv, d, device="cuda", dtype=torch.bfloat16)
    vc = torch.zeros_like(kc)
    slots = torch.randperm(n_pages * page, device="cuda")[:m]
    rows = torch.zeros(64, dtype=torch.int64, device="cuda")
    rows[:m] = slots * (hkv * d)
    ss = torch.rand(64, device="cuda") * k + 1.0
    q = dl.decode_linear(x, wp, dl.EPI_ROPE_APPEND, row_sumsq=ss, norm_dim=k, cos_sin=cs, cache_row=rows, k_cache=kc, v_cache=vc,
                         num_q_heads=hq, num_kv_heads=hkv, head_dim=d, interleave=interleave)
    # oracle: plain composition on the ORIGINAL row order
    h = (x.float() @ w.float().t()) * torch.rsqrt(ss[:m] / k + 1e-5)[:, None]
    hq_, hk_, hv_ = h[:, : hq * d].view(m, hq, d), h[:, hq * d:(hq + hkv) * d].view(m, hkv, d), h[:, (hq + hkv) * d:].view(m, hkv, d)
    c, s_ = torch.cos(ang)[:, None], torch.sin(ang)[:, None]

    def rope(t):
        if interleave:
            x1, x2 = t[..., 0::2], t[..., 1::2]
            return torch.stack([x1 * c - x2 * s_, x2 * c + x1 * s_], -1).flatten(-2)
        x1, x2 = t[..., : d // 2], t[..., d // 2:]
        return torch.cat([x1 * c - x2 * s_, x2 * c + x1 * s_], -1)

    assert _rel(q.view(m, hq, d), rope(hq_)) < 2e-2
    kflat, vflat = kc.view(-1, hkv, d), vc.view(-1, hkv, d)
    assert _rel(kflat[slots], rope(hk_)) < 2e-2
    assert _rel(vflat[slots], hv_) < 2e-2
    untouched = torch.ones(n_pages * page, dtype=torch.bool, device="cuda")
    untouched[slots] = False
    assert float(kflat[untouched].abs().max()) == 0.0 and float(vflat[untouched].abs().max()) == 0.0


def test_decode_prep_matches_cpu():
    m, h, d, page = 64, 1024, 128, 16
    embed = torch.randn(500, h, device="cuda").bfloat16()
    tokens = torch.randint(0, 500, (m,), device="cuda")
    pos = torch.randint(0, 4000, (m,), device="cuda", dtype=torch.int32)
    ppr = 256
    indptr = torch.arange(0, (m + 1) * ppr, ppr, dtype=torch.int32, device="cuda")
    indices = torch.randperm(m * ppr, device="cuda").int()
    outs = []
    for dev in ("cuda", "cpu"):
        res = torch.zeros(m, h, dtype=torch.bfloat16, device=dev)
        ss = torch.ones(5, 64, device=dev)
        cs = torch.zeros(64, d, device=dev)
        row = torch.zeros(64, dtype=torch.int64, device=dev)
        dl.decode_prep(tokens.to(dev), embed.to(dev), res, ss, pos.to(dev), indptr.to(dev), indices.to(dev), page, page * 8 * d, 8 * d,
                       cs, row, d, rope_scale=8.0, rope_theta=5e5, llama31=(1.0, 4.0, 8192.0))
        outs.append((res.cpu(), ss.cpu(), cs.cpu(), row.cpu()))
    (r0, s0, c0, w0), (r1, s1, c1, w1) = outs
    assert torch.equal(r0, r1) and torch.equal(w0, w1)
    torch.testing.assert_close(s0, s1, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(c0, c1, rtol=0, atol=2e-3)  # sin / cos of angles up to 4000 rad in fp32


def test_fused_engine_matches_unfused_gpu():
    from flashinfer_b200.models.llama import LlamaConfig, LlamaDecodeEngine

    cfg = LlamaConfig.tiny()
    cfg.head_dim, cfg.hidden_size = 128, 1024
    batch, kv_len, page = 9, 300, 16
    ppr = (kv_len + page - 1) // page
    indptr = torch.arange(0, (batch + 1) * ppr, ppr, dtype=torch.int32)
    indices = torch.randperm(batch * ppr).int()
    last = torch.full((batch,), (kv_len - 1) % page + 1, dtype=torch.int32)
    logits = []
    for fused in (False, True):
        e = LlamaDecodeEngine(cfg, batch, batch * ppr, page, fused=fused, random_norms=True)
        torch.manual_seed(5)
        for l in e.layers:
            l["k_cache"].copy_(torch.randn_like(l["k_cache"]) * 0.5)
            l["v_cache"].copy_(torch.randn_like(l["v_cache"]) * 0.5)
        e.plan(indptr, indices, last)
        e.tokens.copy_(torch.arange(batch) * 13 % cfg.vocab_size)
        e.step()
        first = e._logits.float().clone()
        e.capture(warmup=1)  # the fused path must be CUDA-graph capturable and replay-stable
        e.replay()
        torch.cuda.synchronize()
        logits.append((first, e._logits.float().clone(), e.layers[0]["k_cache"].float().clone()))
    (u0, u1, uk), (f0, f1, fk) = logits
    assert _rel(f0, u0) < 5e-2 and _rel(f1, u1) < 5e-2
    assert _rel(fk, uk) < 2e-2


# ------------------------------------------------------------------ tensor-parallel epilogue (in-kernel all-reduce)
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _tp_worker(rank, world, port, errs, algo=0):
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        worst = 0.0
        n = 4096
        tp = dl.FusedLinearTP(None, 64, n, torch.bfloat16, algo=algo)
        for it, (m, k, bn, s) in enumerate([(64, 512, 0, 0), (64, 1792, 0, 0), (3, 4096, 0, 0), (64, 2048, 64, 2), (33, 1024, 32, 1), (17, 2048, 128, 4)] * 2):
            torch.manual_seed(100 + it)
            res = torch.randn(m, n, device="cuda").bfloat16()  # replicated residual stream
            torch.manual_seed(7 * it + rank)
            x = (torch.randn(m, k, device="cuda") * 0.5).bfloat16()
            w = (torch.randn(n, k, device="cuda") / k ** 0.5).bfloat16()
            part = (x.float() @ w.float().t()).bfloat16().float()
            dist.all_reduce(part)
            ref = res.float() + part
            ssq = torch.zeros(64, device="cuda")
            dl.decode_linear(x, w, dl.EPI_RESIDUAL, residual=res, sumsq_out=ssq, tp=tp, bn=bn, split_k=s)
            torch.cuda.synchronize()
            worst = max(worst, float((res.float() - ref).abs().max() / ref.abs().max()))
            worst = max(worst, float(((ssq[:m] - res.float().pow(2).sum(-1)).abs() / res.float().pow(2).sum(-1)).max()))
        # CUDA-graph replay of a chain of all-reduce GEMMs (flag epochs must survive replays)
        x = (torch.randn(64, 1024, device="cuda") * 0.5).bfloat16()
        w = (torch.randn(n, 1024, device="cuda") / 32).bfloat16()
        res = torch.zeros(64, n, device="cuda").bfloat16()
        ssq = torch.zeros(64, device="cuda")
        s_ = torch.cuda.Stream()
        with torch.cuda.stream(s_):
            for _ in range(2):
                dl.decode_linear(x, w, dl.EPI_RESIDUAL, residual=res, sumsq_out=ssq, tp=tp)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        res.zero_()
        with torch.cuda.graph(g):
            for _ in range(4):
                dl.decode_linear(x, w, dl.EPI_RESIDUAL, residual=res, sumsq_out=ssq, tp=tp)
        res.zero_()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        part = (x.float() @ w.float().t()).bfloat16().float()
        dist.all_reduce(part)
        worst = max(worst, float((res.float() - 12 * part).abs().max() / (12 * part).abs().max()))
        errs[rank] = worst
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("algo", [1, 2], ids=["one_shot", "two_shot"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_tp_residual_allreduce(world, algo):
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    errs = mp.get_context("spawn").Manager().dict()
    mp.spawn(_tp_worker, args=(world, _free_port(), errs, algo), nprocs=world, join=True)
    assert len(errs) == world and max(errs.values()) < 3e-2, dict(errs)
