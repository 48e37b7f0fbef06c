"""Gated-delta-net decode engine against a plain PyTorch delta-rule recurrence over several steps (slot-addressed states)."""
import torch

from flashinfer_b200.models.gdn import GDNConfig, GDNDecodeEngine


def _rms(x, w, eps):
    x = x.float()
    return x * torch.rsqrt((x * x).mean(-1, keepdim=True) + eps) * w.float()


def test_gdn_engine_matches_plain_recurrence():
    cfg = GDNConfig.tiny()
    eng = GDNDecodeEngine(cfg, max_slots=6, device="cpu", dtype=torch.bfloat16, seed=1)
    slots = torch.tensor([4, 1, 2], dtype=torch.int32)
    eng.plan(slots)
    rd = lambda t: t.to(torch.bfloat16).float()  # noqa: E731
    b, hk, hv, kd, vd = 3, cfg.num_k_heads, cfg.num_v_heads, cfg.head_k_dim, cfg.head_v_dim
    conv = [torch.zeros(b, cfg.qkv_dim, cfg.conv_kernel) for _ in eng.layers]
    state = [torch.zeros(b, hv, kd, vd) for _ in eng.layers]
    g = torch.Generator().manual_seed(9)
    for _ in range(5):
        tok = torch.randint(0, cfg.vocab_size, (b,), generator=g)
        res = eng.embed[tok].float()
        for li, l in enumerate(eng.layers):
            x = rd(_rms(res, l["ln1"], cfg.rms_eps))
            proj = rd(x @ l["in_proj"].float().t())
            qkv, z = proj[:, : cfg.qkv_dim], proj[:, cfg.qkv_dim: cfg.qkv_dim + hv * vd]
            bg, a = proj[:, cfg.qkv_dim + hv * vd: cfg.qkv_dim + hv * vd + hv], proj[:, cfg.qkv_dim + hv * vd + hv:]
            conv[li] = torch.cat([conv[li][:, :, 1:], qkv.unsqueeze(-1)], -1)
            qkv = rd(torch.nn.functional.silu((conv[li] * l["conv_w"].float()).sum(-1)))
            q = qkv[:, : hk * kd].reshape(b, hk, kd).repeat_interleave(hv // hk, 1)
            k = qkv[:, hk * kd: 2 * hk * kd].reshape(b, hk, kd).repeat_interleave(hv // hk, 1)
            v = qkv[:, 2 * hk * kd:].reshape(b, hv, vd)
            q = q * torch.rsqrt((q * q).sum(-1, keepdim=True) + 1e-6)
            k = k * torch.rsqrt((k * k).sum(-1, keepdim=True) + 1e-6)
            gate = torch.exp(-torch.exp(l["A_log"]) * torch.nn.functional.softplus(a + l["dt_bias"]))     # [b, hv]
            beta = torch.sigmoid(bg)
            s = state[li] * gate[..., None, None]
            delta = (v - torch.einsum("bhk,bhkv->bhv", k, s)) * beta[..., None]
            s = s + k[..., None] * delta[:, :, None, :]
            state[li] = s
            o = rd(torch.einsum("bhk,bhkv->bhv", q * kd ** -0.5, s))
            o = rd(_rms(o.reshape(b * hv, vd), l["o_norm"], cfg.rms_eps)).view(b, hv * vd)
            res = rd(res + rd(rd(o * torch.nn.functional.silu(z)) @ l["out_proj"].float().t()))
            x = rd(_rms(res, l["ln2"], cfg.rms_eps))
            gu = rd(x @ l["w_gu"].float().t())
            i = gu.shape[-1] // 2
            res = rd(res + rd(rd(torch.nn.functional.silu(gu[:, :i]) * gu[:, i:]) @ l["w_d"].float().t()))
        want = rd(rd(_rms(res, eng.final_norm, cfg.rms_eps)) @ eng.lm_head.float().t())
        eng.tokens.copy_(tok)
        eng.step()
        got = eng.logits.float()
        cos = torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0)
        assert cos > 0.999, float(cos)
        torch.testing.assert_close(got, want, atol=0.05 * float(want.abs().max()), rtol=0.05)
    torch.testing.assert_close(eng.layers[0]["state"][slots.long()], state[0], atol=2e-2, rtol=2e-2)
    assert float(eng.layers[0]["state"][[0, 3, 5]].abs().sum()) == 0.0
