"""Mamba-2 decode engine against a plain PyTorch recurrence (conv window, selective scan step, gated norm) over several steps,
with batch rows bound to non-contiguous state slots."""
import torch

from flashinfer_b200.models.mamba2 import Mamba2Config, Mamba2DecodeEngine


def _rms(x, w, eps):
    x = x.float()
    return x * torch.rsqrt((x * x).mean(-1, keepdim=True) + eps) * w.float()


class _Reference:
    def __init__(self, eng, batch):
        self.eng = eng
        cfg = eng.cfg
        self.conv = [torch.zeros(batch, cfg.conv_dim, cfg.conv_kernel) for _ in eng.layers]
        self.ssm = [torch.zeros(batch, cfg.num_heads, cfg.head_dim, cfg.state_size) for _ in eng.layers]

    def step(self, tokens):
        eng, cfg = self.eng, self.eng.cfg
        rd = lambda t: t.to(eng.dtype).float()  # noqa: E731
        b, hn, p, n, g = tokens.numel(), cfg.num_heads, cfg.head_dim, cfg.state_size, cfg.n_groups
        res = eng.embed[tokens].float()
        for li, l in enumerate(eng.layers):
            x = rd(_rms(res, l["ln"], cfg.rms_eps))
            proj = rd(x @ l["in_proj"].float().t())
            z, xbc, dt = proj[:, : cfg.d_inner], proj[:, cfg.d_inner: cfg.d_inner + cfg.conv_dim], proj[:, cfg.d_inner + cfg.conv_dim:]
            self.conv[li] = torch.cat([self.conv[li][:, :, 1:], xbc.unsqueeze(-1)], -1)
            xbc = rd(torch.nn.functional.silu((self.conv[li] * l["conv_w"].float()).sum(-1) + l["conv_b"].float()))
            xs = xbc[:, : cfg.d_inner].reshape(b, hn, p)
            bm = xbc[:, cfg.d_inner: cfg.d_inner + g * n].reshape(b, g, n).repeat_interleave(hn // g, 1)
            cm = xbc[:, cfg.d_inner + g * n:].reshape(b, g, n).repeat_interleave(hn // g, 1)
            step = torch.nn.functional.softplus(dt[:, :, None] + l["dt_bias"])                      # [b, H, P]
            self.ssm[li] = self.ssm[li] * torch.exp(l["A"] * step[..., None]) + (step * xs)[..., None] * bm[:, :, None, :]
            y = (self.ssm[li] * cm[:, :, None, :]).sum(-1) + l["D"] * xs
            y = rd(y * torch.nn.functional.silu(z.reshape(b, hn, p)))
            res = rd(res + rd(rd(_rms(y.reshape(b, -1), l["gate_norm"], cfg.rms_eps)) @ l["out_proj"].float().t()))
        return rd(rd(_rms(res, eng.final_norm, cfg.rms_eps)) @ eng.lm_head.float().t())


def test_mamba2_engine_matches_plain_recurrence():
    cfg = Mamba2Config.tiny()
    eng = Mamba2DecodeEngine(cfg, max_slots=7, device="cpu", dtype=torch.bfloat16, seed=2)
    slots = torch.tensor([5, 0, 3], dtype=torch.int32)
    eng.plan(slots)
    ref = _Reference(eng, 3)
    g = torch.Generator().manual_seed(4)
    for _ in range(5):                                                      # the conv window fills and the state accumulates
        tok = torch.randint(0, cfg.vocab_size, (3,), generator=g)
        eng.tokens.copy_(tok)
        want = ref.step(tok)
        eng.step()
        got = eng.logits.float()
        cos = torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0)
        assert cos > 0.999, float(cos)
        torch.testing.assert_close(got, want, atol=0.05 * float(want.abs().max()), rtol=0.05)
    untouched = [s for s in range(7) if s not in slots.tolist()]
    assert float(eng.layers[0]["ssm_state"][untouched].abs().sum()) == 0.0 and float(eng.layers[0]["ssm_state"][slots.long()].abs().sum()) > 0
    torch.testing.assert_close(eng.layers[1]["ssm_state"][slots.long()], ref.ssm[1], atol=2e-2, rtol=2e-2)
    assert cfg.conv_dim == cfg.d_inner + 2 * cfg.state_size and Mamba2Config.mamba2_2_7b().num_heads == 80
