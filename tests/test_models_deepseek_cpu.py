"""DeepSeek-style engine (absorbed MLA + grouped-top-k MoE with a shared expert) against a plain, NON-absorbed PyTorch model that
expands keys / values per head from the latent cache - checks the weight absorption, the interleaved RoPE, the latent-cache append,
the routing and the expert / shared-expert combination end to end through the public ops (CPU eager paths)."""
import math

import pytest
import torch

from flashinfer_b200.models.deepseek import DeepSeekConfig, DeepSeekDecodeEngine


def _rms(x, w, eps):
    x = x.float()
    return x * torch.rsqrt((x * x).mean(-1, keepdim=True) + eps) * w.float()


def _rope_interleaved(x, pos, theta):
    """x [..., d] rotated in (even, odd) pairs by angle pos * theta^(-2 i / d)."""
    d = x.shape[-1]
    inv = theta ** (-torch.arange(0, d, 2).float() / d)
    ang = pos.float()[..., None] * inv
    cos, sin = ang.cos(), ang.sin()
    xe, xo = x[..., 0::2].float(), x[..., 1::2].float()
    out = torch.empty_like(x, dtype=torch.float32)
    out[..., 0::2] = xe * cos - xo * sin
    out[..., 1::2] = xo * cos + xe * sin
    return out


def _reference_step(eng, tokens, history):
    """history[layer][request] = (ckv [n - 1, rank], kpe [n - 1, rope]) already in the cache; returns logits [b, vocab] in fp32."""
    cfg = eng.cfg
    b, hq = tokens.numel(), cfg.num_heads
    h = eng.embed[tokens].float()
    for li, l in enumerate(eng.layers):
        x = _rms(h, l["ln1"], cfg.rms_eps).to(eng.dtype)
        if cfg.q_lora_rank:
            q = _rms((x.float() @ l["q_a"].float().t()).to(eng.dtype), l["q_norm"], cfg.rms_eps).to(eng.dtype).float() @ l["q_b"].float().t()
        else:
            q = x.float() @ l["q_proj"].float().t()
        q = q.to(eng.dtype).float().view(b, hq, cfg.qk_head_dim)
        kv = (x.float() @ l["kv_a"].float().t()).to(eng.dtype)
        ckv_new = _rms(kv[:, : cfg.kv_lora_rank], l["kv_norm"], cfg.rms_eps).to(eng.dtype).float()
        attn = torch.zeros(b, hq, cfg.v_head_dim)
        for r in range(b):
            ckv_old, kpe_old = history[li][r]
            pos = ckv_old.shape[0]
            q_pe = _rope_interleaved(q[r, :, cfg.qk_nope_head_dim:], torch.full((hq,), pos), cfg.rope_theta).to(eng.dtype).float()
            k_pe = _rope_interleaved(kv[r, cfg.kv_lora_rank:].float()[None], torch.tensor([pos]), cfg.rope_theta).to(eng.dtype).float()
            ckv = torch.cat([ckv_old.float(), ckv_new[r][None]])                      # [n, rank]
            kpe = torch.cat([kpe_old.float(), k_pe])                                   # [n, rope]
            k_nope = torch.einsum("hdr,nr->nhd", l["w_uk"].float(), ckv)               # keys expanded per head
            val = torch.einsum("hdr,nr->nhd", l["w_uv"].float(), ckv)                  # values expanded per head
            logit = (torch.einsum("hd,nhd->hn", q[r, :, : cfg.qk_nope_head_dim], k_nope) + q_pe @ kpe.t()) * cfg.softmax_scale
            attn[r] = torch.einsum("hn,nhd->hd", torch.softmax(logit, -1), val)
        h = h + attn.reshape(b, -1) @ l["wo"].float().t()
        h = h.to(eng.dtype).float()
        x = _rms(h, l["ln2"], cfg.rms_eps).to(eng.dtype).float()
        if "w_gu" in l:
            gu = x @ l["w_gu"].float().t()
            inter = gu.shape[-1] // 2
            f = (torch.nn.functional.silu(gu[:, :inter]) * gu[:, inter:]) @ l["w_d"].float().t()
        else:
            s = torch.sigmoid(x @ l["router"].t())
            sb = s + l["router_bias"]
            e = cfg.num_experts
            grp = sb.view(b, cfg.n_group, e // cfg.n_group)
            gscore = grp.topk(2, -1).values.sum(-1)
            keep = torch.zeros_like(gscore, dtype=torch.bool).scatter_(1, gscore.topk(cfg.topk_group, -1).indices, True)
            masked = torch.where(keep[..., None].expand_as(grp).reshape(b, e), sb, torch.full_like(sb, float("-inf")))
            ids = masked.topk(cfg.num_experts_per_tok, -1).indices
            wts = s.gather(1, ids)
            wts = wts / wts.sum(-1, keepdim=True) * cfg.routed_scaling_factor
            i = cfg.moe_intermediate_size
            f = torch.zeros(b, cfg.hidden_size)
            for r in range(b):
                for j in range(cfg.num_experts_per_tok):
                    ex = int(ids[r, j])
                    hid = l["w1"][ex].float() @ x[r]
                    f[r] += wts[r, j] * (l["w2"][ex].float() @ (torch.nn.functional.silu(hid[i:]) * hid[:i]))
            sg = x @ l["shared_gu"].float().t()
            si = sg.shape[-1] // 2
            f = f + (torch.nn.functional.silu(sg[:, :si]) * sg[:, si:]) @ l["shared_d"].float().t()
        h = (h + f).to(eng.dtype).float()
    return _rms(h, eng.final_norm, cfg.rms_eps).to(eng.dtype).float() @ eng.lm_head.float().t()


@pytest.mark.parametrize("q_lora", [True, False])
def test_deepseek_engine_matches_non_absorbed_model(q_lora):
    cfg = DeepSeekConfig.tiny()
    if not q_lora:
        cfg.q_lora_rank = None
    page_size, lens = 8, [13, 1, 24]                              # lengths INCLUDE the token decoded in this step
    per = [(n + page_size - 1) // page_size for n in lens]
    g = torch.Generator().manual_seed(7)
    ids = torch.randperm(sum(per) + 2, generator=g)[: sum(per)].int()
    indptr = torch.tensor([0] + list(torch.tensor(per).cumsum(0)), dtype=torch.int32)
    last = torch.tensor([(n - 1) % page_size + 1 for n in lens], dtype=torch.int32)
    eng = DeepSeekDecodeEngine(cfg, max_batch=4, max_pages=sum(per) + 2, page_size=page_size, device="cpu", dtype=torch.bfloat16, seed=3)
    history = []
    for l in eng.layers:                                           # random history in the latent cache (positions 0 .. n - 2)
        l["ckv_cache"].copy_((torch.randn(l["ckv_cache"].shape, generator=g) * 0.5).to(torch.bfloat16))
        l["kpe_cache"].copy_((torch.randn(l["kpe_cache"].shape, generator=g) * 0.5).to(torch.bfloat16))
        per_req = []
        for r, n in enumerate(lens):
            pages = ids[int(indptr[r]): int(indptr[r + 1])].long()
            per_req.append((l["ckv_cache"][pages].reshape(-1, cfg.kv_lora_rank)[: n - 1].clone(),
                            l["kpe_cache"][pages].reshape(-1, cfg.qk_rope_head_dim)[: n - 1].clone()))
        history.append(per_req)
    eng.plan(indptr, ids, last)
    eng.tokens.copy_(torch.randint(0, cfg.vocab_size, (3,), generator=g))
    ref = _reference_step(eng, eng.tokens.clone(), history)
    nxt = eng.step()
    got = eng.logits.float()
    cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0)
    assert cos > 0.999, float(cos)
    torch.testing.assert_close(got, ref, atol=0.06 * float(ref.abs().max()), rtol=0.05)
    assert nxt.shape == (3,) and int(nxt.max()) < cfg.vocab_size
    # the step appended (ckv, k_pe) of the new token at position n - 1 of every request's last page
    for r, n in enumerate(lens):
        pg = int(ids[int(indptr[r + 1]) - 1])
        assert float(eng.layers[0]["ckv_cache"][pg, (n - 1) % page_size].float().abs().sum()) > 0
    # a second step on the grown cache runs (positions advance by re-planning)
    lens2 = [n + 1 for n in lens]
    if all((n + page_size - 1) // page_size == p for n, p in zip(lens2, per)):
        eng.plan(indptr, ids, torch.tensor([(n - 1) % page_size + 1 for n in lens2], dtype=torch.int32))
        eng.tokens.copy_(nxt)
        assert eng.step().shape == (3,)


def test_config_presets():
    v3 = DeepSeekConfig.deepseek_v3()
    assert v3.qk_head_dim == 192 and abs(v3.softmax_scale - 1 / math.sqrt(192)) < 1e-9 and v3.num_experts % v3.n_group == 0
    t = DeepSeekConfig.tiny()
    assert t.kv_lora_rank == 512 and t.qk_rope_head_dim == 64 and t.first_k_dense < t.num_layers


def test_prefill_equals_token_by_token_decode():
    """One causal multi-token pass over the latent cache gives the logits of feeding the same tokens one decode step at a time."""
    cfg = DeepSeekConfig.tiny()
    page_size, prompts = 8, [[5, 9, 2, 77, 41, 3, 8, 120, 6, 1], [300, 4, 18]]
    pages = [[0, 3], [2]]
    mk = lambda: DeepSeekDecodeEngine(cfg, max_batch=2, max_pages=5, page_size=page_size, device="cpu", dtype=torch.bfloat16, seed=9)  # noqa: E731

    def tables(lens):
        used = [pages[r][: -(-n // page_size)] for r, n in enumerate(lens)]
        indptr = torch.tensor([0, len(used[0]), len(used[0]) + len(used[1])], dtype=torch.int32)
        return indptr, torch.tensor(used[0] + used[1], dtype=torch.int32), torch.tensor([(n - 1) % page_size + 1 for n in lens], dtype=torch.int32)

    a = mk()
    lens = [len(p) for p in prompts]
    a.prefill(torch.tensor(prompts[0] + prompts[1]), torch.tensor([0, lens[0], lens[0] + lens[1]], dtype=torch.int32), *tables(lens), all_logits=True)
    all_rows = a.logits.float().clone()
    b = mk()
    for r, p in enumerate(prompts):                                  # request by request, token by token
        for i, tok in enumerate(p):
            cur = [0, 0]
            cur[r] = i + 1
            indptr, idx, last = tables([max(cur[0], 1), max(cur[1], 1)])
            sel = slice(int(indptr[r]), int(indptr[r + 1]))
            b.plan(torch.tensor([0, -(-(i + 1) // page_size)], dtype=torch.int32), idx[sel], last[r:r + 1])
            b.tokens.copy_(torch.tensor([tok]))
            b.step()
            row = (0 if r == 0 else lens[0]) + i
            cos = torch.nn.functional.cosine_similarity(b.logits[0].float(), all_rows[row], dim=0)
            assert cos > 0.999, (r, i, float(cos))


def test_generate_with_the_mla_engine():
    from flashinfer_b200.models.serving import generate

    cfg = DeepSeekConfig.tiny()
    prompts = [[5, 9, 2, 77, 41], [300, 4]]
    mk = lambda b: DeepSeekDecodeEngine(cfg, max_batch=b, max_pages=8, page_size=8, device="cpu", dtype=torch.bfloat16, seed=9)  # noqa: E731
    batched = generate(mk(2), prompts, 4)
    assert [len(o) for o in batched] == [4, 4]
    assert generate(mk(1), [prompts[1]], 4)[0] == batched[1]             # batching does not change a request's tokens
