"""The configurable GQA decode engine (Mixtral / Qwen3 / Qwen3-MoE / Gemma-2 switches) against a plain PyTorch implementation of
the same architecture options: q/k norm, (1 + w) norms, post norms, GeGLU, embedding scale, soft-caps, sliding-window layers,
renormalised top-k MoE routing."""
import math

import pytest
import torch

from flashinfer_b200.models.transformer import TransformerConfig, TransformerDecodeEngine


def _norm(x, w, eps, gemma):
    x = x.float()
    return x * torch.rsqrt((x * x).mean(-1, keepdim=True) + eps) * (w.float() + (1.0 if gemma else 0.0))


def _rope_neox(x, pos, theta):
    d = x.shape[-1]
    inv = theta ** (-torch.arange(0, d, 2).float() / d)
    ang = pos.float()[..., None] * inv
    cos, sin = ang.cos(), ang.sin()
    x1, x2 = x[..., : d // 2].float(), x[..., d // 2:].float()
    return torch.cat([x1 * cos - x2 * sin, x2 * cos + x1 * sin], -1)


def _reference_step(eng, tokens, history):
    cfg, dt = eng.cfg, eng.dtype
    rd = lambda t: t.to(dt).float()  # noqa: E731  (round like the engine's bf16 activations)
    b, hq, hkv, d = tokens.numel(), cfg.num_qo_heads, cfg.num_kv_heads, cfg.head_dim
    res = rd(eng.embed[tokens].float() * cfg.embed_scale)
    for li, l in enumerate(eng.layers):
        x = rd(_norm(res, l["ln1"], cfg.rms_eps, cfg.gemma_norm))
        qkv = rd(x @ l["wqkv"].float().t()).view(b, hq + 2 * hkv, d)
        q, k, v = qkv[:, :hq], qkv[:, hq:hq + hkv], qkv[:, hq + hkv:]
        if cfg.qk_norm:
            q, k = rd(_norm(q, l["q_norm"], cfg.rms_eps, cfg.gemma_norm)), rd(_norm(k, l["k_norm"], cfg.rms_eps, cfg.gemma_norm))
        attn = torch.zeros(b, hq, d)
        for r in range(b):
            k_old, v_old = history[li][r]
            pos = k_old.shape[0]
            qr = rd(_rope_neox(q[r], torch.full((hq,), pos), cfg.rope_theta))
            kr = rd(_rope_neox(k[r], torch.full((hkv,), pos), cfg.rope_theta))
            keys, vals = torch.cat([k_old.float(), kr[None]]), torch.cat([v_old.float(), v[r][None]])         # [n, hkv, d]
            if cfg.is_sliding(li):
                keys, vals = keys[-cfg.sliding_window:], vals[-cfg.sliding_window:]
            keys, vals = keys.repeat_interleave(hq // hkv, 1), vals.repeat_interleave(hq // hkv, 1)
            lg = torch.einsum("hd,nhd->hn", qr, keys) * cfg.softmax_scale
            if cfg.attn_logit_softcap:
                lg = cfg.attn_logit_softcap * torch.tanh(lg / cfg.attn_logit_softcap)
            attn[r] = torch.einsum("hn,nhd->hd", torch.softmax(lg, -1), vals)
        a = rd(rd(attn).reshape(b, -1) @ l["wo"].float().t())
        if cfg.post_norms:
            a = rd(_norm(a, l["post_attn"], cfg.rms_eps, cfg.gemma_norm))
        res = rd(res + a)
        x = rd(_norm(res, l["ln2"], cfg.rms_eps, cfg.gemma_norm))
        if cfg.num_experts:
            lg = rd(x @ l["router"].float().t())
            top, ids = lg.topk(cfg.num_experts_per_tok, -1)
            wts = torch.softmax(top, -1)
            i = cfg.intermediate_size
            f = torch.zeros(b, cfg.hidden_size)
            for r in range(b):
                for j in range(cfg.num_experts_per_tok):
                    hid = l["w1"][int(ids[r, j])].float() @ x[r]
                    f[r] += wts[r, j] * (l["w2"][int(ids[r, j])].float() @ (torch.nn.functional.silu(hid[i:]) * hid[:i]))
        else:
            gu = rd(x @ l["w_gu"].float().t())
            i = gu.shape[-1] // 2
            act = torch.nn.functional.silu(gu[:, :i]) if cfg.activation == "silu" else torch.nn.functional.gelu(gu[:, :i], approximate="tanh")
            f = rd(act * gu[:, i:]) @ l["w_d"].float().t()
        f = rd(f)
        if cfg.post_norms:
            f = rd(_norm(f, l["post_ffn"], cfg.rms_eps, cfg.gemma_norm))
        res = rd(res + f)
    logits = rd(rd(_norm(res, eng.final_norm, cfg.rms_eps, cfg.gemma_norm)) @ eng.lm_head.float().t())
    if cfg.final_logit_softcap:
        logits = torch.tanh(logits / cfg.final_logit_softcap) * cfg.final_logit_softcap
    return logits


@pytest.mark.parametrize("preset", ["mixtral_8x7b", "qwen3_8b", "qwen3_30b_a3b", "gemma2_9b", "plain"])
def test_engine_matches_plain_pytorch(preset):
    cfg = (TransformerConfig() if preset == "plain" else getattr(TransformerConfig, preset)()).tiny()
    page_size, lens = 4, [9, 1, 15]                                   # longer than the tiny sliding window (6)
    per = [(n + page_size - 1) // page_size for n in lens]
    g = torch.Generator().manual_seed(11)
    ids = torch.randperm(sum(per) + 2, generator=g)[: sum(per)].int()
    indptr = torch.tensor([0] + list(torch.tensor(per).cumsum(0)), dtype=torch.int32)
    last = torch.tensor([(n - 1) % page_size + 1 for n in lens], dtype=torch.int32)
    eng = TransformerDecodeEngine(cfg, max_batch=4, max_pages=sum(per) + 2, page_size=page_size, device="cpu", dtype=torch.bfloat16, seed=5)
    history = []
    for l in eng.layers:
        for name in ("k_cache", "v_cache"):
            l[name].copy_((torch.randn(l[name].shape, generator=g) * 0.5).to(torch.bfloat16))
        per_req = []
        for r, n in enumerate(lens):
            pages = ids[int(indptr[r]): int(indptr[r + 1])].long()
            per_req.append(tuple(l[name][pages].reshape(-1, cfg.num_kv_heads, cfg.head_dim)[: n - 1].clone() for name in ("k_cache", "v_cache")))
        history.append(per_req)
    eng.plan(indptr, ids, last)
    eng.tokens.copy_(torch.randint(0, cfg.vocab_size, (3,), generator=g))
    ref = _reference_step(eng, eng.tokens.clone(), history)
    nxt = eng.step()
    got = eng.logits.float()
    cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0)
    assert cos > 0.998, float(cos)
    torch.testing.assert_close(got, ref, atol=0.08 * float(ref.abs().max()), rtol=0.05)
    assert nxt.shape == (3,)


def test_presets_and_layer_pattern():
    g2 = TransformerConfig.gemma2_9b()
    assert g2.head_dim == 256 and [g2.is_sliding(i) for i in range(4)] == [True, False, True, False] and abs(g2.softmax_scale - 1 / 16) < 1e-9
    assert not TransformerConfig.mixtral_8x7b().is_sliding(0) and TransformerConfig.qwen3_30b_a3b().num_experts == 128
    t = g2.tiny()
    assert t.sliding_window == 6 and t.embed_scale == math.sqrt(t.hidden_size) and t.name.endswith("-tiny")


def _full_forward_logits(eng, seq):
    """Plain causal forward of a whole token sequence (no cache): logits of the LAST position.  Reuses _reference_step token by
    token with an explicit per-layer K / V history (keys stored after RoPE, like the cache)."""
    cfg = eng.cfg
    hist = [[(torch.zeros(0, cfg.num_kv_heads, cfg.head_dim), torch.zeros(0, cfg.num_kv_heads, cfg.head_dim))] for _ in eng.layers]
    logits = None
    for pos, tok in enumerate(seq):
        # run one step and capture the k / v this step appends, by replaying the projection of the reference
        logits = _reference_step(eng, torch.tensor([tok]), hist)
        dt = eng.dtype
        rd = lambda t: t.to(dt).float()  # noqa: E731
        res = rd(eng.embed[torch.tensor([tok])].float() * cfg.embed_scale)
        new_hist = []
        for li, l in enumerate(eng.layers):
            x = rd(_norm(res, l["ln1"], cfg.rms_eps, cfg.gemma_norm))
            hq, hkv, d = cfg.num_qo_heads, cfg.num_kv_heads, cfg.head_dim
            qkv = rd(x @ l["wqkv"].float().t()).view(1, hq + 2 * hkv, d)
            k, v = qkv[:, hq:hq + hkv], qkv[:, hq + hkv:]
            if cfg.qk_norm:
                k = rd(_norm(k, l["k_norm"], cfg.rms_eps, cfg.gemma_norm))
            kr = rd(_rope_neox(k[0], torch.full((hkv,), pos), cfg.rope_theta))
            k_old, v_old = hist[li][0]
            new_hist.append([(torch.cat([k_old, kr[None]]), torch.cat([v_old, v[0][None]]))])
            # advance the residual exactly like _reference_step does (recompute this layer)
            single = _LayerOnly(eng, li)
            res = single(res, hist[li][0], pos)
        hist = new_hist
    return logits


class _LayerOnly:
    """One layer of the plain model on one token (used to advance the residual while collecting K / V history)."""

    def __init__(self, eng, li):
        self.eng, self.li = eng, li

    def __call__(self, res, kv_hist, pos):
        eng, cfg, l = self.eng, self.eng.cfg, self.eng.layers[self.li]
        sub = type("E", (), {})()
        sub.cfg, sub.dtype, sub.embed, sub.final_norm, sub.lm_head = cfg, eng.dtype, eng.embed, eng.final_norm, eng.lm_head
        sub.layers = [l]
        # run the single layer through the same reference code path: feed the residual as the "embedding" of a fake token
        rd = lambda t: t.to(eng.dtype).float()  # noqa: E731
        hq, hkv, d = cfg.num_qo_heads, cfg.num_kv_heads, cfg.head_dim
        x = rd(_norm(res, l["ln1"], cfg.rms_eps, cfg.gemma_norm))
        qkv = rd(x @ l["wqkv"].float().t()).view(1, hq + 2 * hkv, d)
        q, k, v = qkv[:, :hq], qkv[:, hq:hq + hkv], qkv[:, hq + hkv:]
        if cfg.qk_norm:
            q, k = rd(_norm(q, l["q_norm"], cfg.rms_eps, cfg.gemma_norm)), rd(_norm(k, l["k_norm"], cfg.rms_eps, cfg.gemma_norm))
        qr = rd(_rope_neox(q[0], torch.full((hq,), pos), cfg.rope_theta))
        kr = rd(_rope_neox(k[0], torch.full((hkv,), pos), cfg.rope_theta))
        keys, vals = torch.cat([kv_hist[0], kr[None]]), torch.cat([kv_hist[1], v[0][None]])
        if cfg.is_sliding(self.li):
            keys, vals = keys[-cfg.sliding_window:], vals[-cfg.sliding_window:]
        keys, vals = keys.repeat_interleave(hq // hkv, 1), vals.repeat_interleave(hq // hkv, 1)
        lg = torch.einsum("hd,nhd->hn", qr, keys) * cfg.softmax_scale
        if cfg.attn_logit_softcap:
            lg = cfg.attn_logit_softcap * torch.tanh(lg / cfg.attn_logit_softcap)
        a = rd(rd(torch.einsum("hn,nhd->hd", torch.softmax(lg, -1), vals)).reshape(1, -1) @ l["wo"].float().t())
        if cfg.post_norms:
            a = rd(_norm(a, l["post_attn"], cfg.rms_eps, cfg.gemma_norm))
        res = rd(res + a)
        x = rd(_norm(res, l["ln2"], cfg.rms_eps, cfg.gemma_norm))
        if cfg.num_experts:
            lgts = rd(x @ l["router"].float().t())
            top, ids = lgts.topk(cfg.num_experts_per_tok, -1)
            wts = torch.softmax(top, -1)
            i = cfg.intermediate_size
            f = torch.zeros(1, cfg.hidden_size)
            for j in range(cfg.num_experts_per_tok):
                hid = l["w1"][int(ids[0, j])].float() @ x[0]
                f[0] += wts[0, j] * (l["w2"][int(ids[0, j])].float() @ (torch.nn.functional.silu(hid[i:]) * hid[:i]))
        else:
            gu = rd(x @ l["w_gu"].float().t())
            i = gu.shape[-1] // 2
            act = torch.nn.functional.silu(gu[:, :i]) if cfg.activation == "silu" else torch.nn.functional.gelu(gu[:, :i], approximate="tanh")
            f = rd(act * gu[:, i:]) @ l["w_d"].float().t()
        f = rd(f)
        if cfg.post_norms:
            f = rd(_norm(f, l["post_ffn"], cfg.rms_eps, cfg.gemma_norm))
        return rd(res + f)


@pytest.mark.parametrize("preset", ["plain", "gemma2_9b", "qwen3_30b_a3b"])
def test_prefill_then_decode_equals_full_sequence_forward(preset):
    """Serve a request end to end: prefill a prompt into the paged cache, then decode two tokens; the logits after every stage equal a
    plain causal forward over the whole sequence so far (sliding-window layers included)."""
    cfg = (TransformerConfig() if preset == "plain" else getattr(TransformerConfig, preset)()).tiny()
    page_size = 4
    prompts = [[5, 17, 3, 99, 42, 7, 250, 11, 8], [200, 1, 64]]
    g = torch.Generator().manual_seed(3)
    max_pages = 12
    eng = TransformerDecodeEngine(cfg, max_batch=2, max_pages=max_pages, page_size=page_size, device="cpu", dtype=torch.bfloat16, seed=8)
    free = torch.randperm(max_pages, generator=g).tolist()
    pages = [[], []]

    def tables(lens):
        for r, n in enumerate(lens):
            while len(pages[r]) * page_size < n:
                pages[r].append(free.pop())
        indptr = torch.tensor([0, len(pages[0]), len(pages[0]) + len(pages[1])], dtype=torch.int32)
        indices = torch.tensor(pages[0] + pages[1], dtype=torch.int32)
        last = torch.tensor([(n - 1) % page_size + 1 for n in lens], dtype=torch.int32)
        return indptr, indices, last

    lens = [len(p) for p in prompts]
    qo = torch.tensor([0, lens[0], lens[0] + lens[1]], dtype=torch.int32)
    nxt = eng.prefill(torch.tensor(prompts[0] + prompts[1]), qo, *tables(lens))
    seqs = [list(p) for p in prompts]
    for r in range(2):
        want = _full_forward_logits(eng, seqs[r])[0]
        got = eng.logits[r].float()
        assert torch.nn.functional.cosine_similarity(got, want, dim=0) > 0.998, (preset, "prefill", r)
    for _ in range(2):
        for r in range(2):
            seqs[r].append(int(nxt[r]))
        lens = [len(s) for s in seqs]
        eng.plan(*tables(lens))
        eng.tokens.copy_(torch.tensor([s[-1] for s in seqs]))
        nxt = eng.step().clone()
        for r in range(2):
            want = _full_forward_logits(eng, seqs[r])[0]
            assert torch.nn.functional.cosine_similarity(eng.logits[r].float(), want, dim=0) > 0.998, (preset, "decode", r)


def test_generation_loop_and_page_allocator():
    """models.serving: batched prefill + greedy decode with on-demand page growth gives, per request, the same tokens as serving the
    request alone; pages return to the pool."""
    from flashinfer_b200.models.serving import PagedKVAllocator, generate

    cfg = TransformerConfig.qwen3_8b().tiny()
    prompts = [[3, 14, 15, 92, 65, 35, 89], [79, 32], [38, 46, 26, 43, 38]]
    eng = TransformerDecodeEngine(cfg, max_batch=3, max_pages=24, page_size=4, device="cpu", dtype=torch.bfloat16, seed=4)
    alloc = PagedKVAllocator(24, 4, seed=1)
    batched = generate(eng, prompts, 5, alloc)
    assert [len(o) for o in batched] == [5, 5, 5] and alloc.free_pages == 24 - sum(-(-(len(p) + 4) // 4) for p in prompts)
    for i, p in enumerate(prompts):
        solo = TransformerDecodeEngine(cfg, max_batch=1, max_pages=24, page_size=4, device="cpu", dtype=torch.bfloat16, seed=4)
        assert generate(solo, [p], 5)[0] == batched[i]                       # batching and page placement do not change the result
    a = PagedKVAllocator(4, 4)
    a.add_request(0)
    a.grow(0, 9)
    assert a.free_pages == 1 and a.tables([0])[2].tolist() == [1] and a.length(0) == 9
    with pytest.raises(MemoryError):
        a.grow(0, 8)
    a.release(0)
    assert a.free_pages == 4
    with pytest.raises(KeyError):
        a.add_request(1) or a.add_request(1)


@pytest.mark.parametrize("same_draft", [True, False])
def test_speculative_decoding_reproduces_greedy_target(same_draft):
    """models.serving.speculative_generate: at temperature 0 the accepted + corrected tokens are exactly the target's greedy tokens for
    ANY draft model; with the target as its own draft every proposal is accepted (k + 1 tokens per round)."""
    from flashinfer_b200.models.serving import generate, speculative_generate

    cfg = TransformerConfig.qwen3_8b().tiny()
    prompts = [[3, 14, 15, 92, 65], [35, 89, 79]]
    mk = lambda seed: TransformerDecodeEngine(cfg, max_batch=2, max_pages=32, page_size=4, device="cpu", dtype=torch.bfloat16, seed=seed)  # noqa: E731
    want = generate(mk(4), prompts, 9)
    got, per_round = speculative_generate(mk(4), mk(4 if same_draft else 5), prompts, 9, num_draft_tokens=3, temperature=0.0)
    assert got == want
    assert per_round == 4.0 if same_draft else 1.0 <= per_round <= 4.0
    sampled, rate = speculative_generate(mk(4), mk(5), prompts, 6, num_draft_tokens=2, temperature=0.8, generator=torch.Generator().manual_seed(0))
    assert all(len(o) == 6 and all(0 <= t < cfg.vocab_size for t in o) for o in sampled) and 1.0 <= rate <= 3.0
