"""Fused decode path (gemm/decode_linear.py + LlamaDecodeEngine(fused=True)) against the op-by-op composition.

CPU tensors run the fp32 oracles of every fused epilogue, so this checks the ALGEBRA of the fusion (RMSNorm folding, RoPE row
permutation, sum-of-squares hand-over between GEMMs, in-epilogue paged-KV append); the CUDA kernels are checked against the same
oracles in tests/test_gpu_decode_linear.py."""
import pytest
import torch

from flashinfer_b200.gemm import decode_linear as dl
from flashinfer_b200.models.llama import LlamaConfig, LlamaDecodeEngine


def _engines(dtype, **kw):
    cfg = LlamaConfig.tiny()
    cfg.head_dim = 64
    batch, kv_len, page = 5, 37, 4
    ppr = (kv_len + page - 1) // page
    indptr = torch.arange(0, (batch + 1) * ppr, ppr, dtype=torch.int32)
    indices = torch.randperm(batch * ppr, generator=torch.Generator().manual_seed(3)).int()
    last = torch.full((batch,), (kv_len - 1) % page + 1, dtype=torch.int32)
    engs = []
    for fused in (False, True):
        e = LlamaDecodeEngine(cfg, batch, batch * ppr, page, device="cpu", dtype=dtype, fused=fused, random_norms=True, **kw)
        torch.manual_seed(11)
        for l in e.layers:
            l["k_cache"].copy_(torch.randn_like(l["k_cache"]) * 0.5)
            l["v_cache"].copy_(torch.randn_like(l["v_cache"]) * 0.5)
        e.plan(indptr, indices, last)
        e.tokens.copy_(torch.arange(batch) * 7 % cfg.vocab_size)
        engs.append(e)
    return engs


def test_fused_step_matches_unfused_fp32():
    ref, fus = _engines(torch.float32)
    ref.step()
    fus.step()
    torch.testing.assert_close(fus._logits, ref._logits, rtol=2e-4, atol=2e-4)
    assert torch.equal(fus.next_tokens, ref.next_tokens)
    # the QKV epilogue wrote this step's K / V into the same cache slots as rope + append_paged_kv_cache
    for lr, lf in zip(ref.layers, fus.layers):
        torch.testing.assert_close(lf["k_cache"], lr["k_cache"], rtol=2e-4, atol=2e-4)
        torch.testing.assert_close(lf["v_cache"], lr["v_cache"], rtol=2e-4, atol=2e-4)


def test_fused_step_matches_unfused_bf16():
    ref, fus = _engines(torch.bfloat16)
    ref.step()
    fus.step()
    torch.testing.assert_close(fus._logits.float(), ref._logits.float(), rtol=6e-2, atol=6e-2)


def test_permute_rope_rows_is_a_permutation():
    w = torch.arange(4 * 8 * 3, dtype=torch.float32).view(32, 3)  # hq=2, hkv=1, d=8 -> (2 + 2*1) * 8 rows
    p = dl.permute_rope_rows(w, 2, 1, 8)
    assert sorted(p[:, 0].tolist()) == sorted(w[:, 0].tolist())
    head0 = p[:8, 0] / 3
    assert head0.tolist() == [0, 4, 1, 5, 2, 6, 3, 7]
    assert torch.equal(p[24:], w[24:])  # V rows untouched


def test_decode_prep_cpu():
    m, h, d, page = 3, 16, 8, 4
    embed = torch.randn(10, h)
    res = torch.zeros(m, h)
    ss = torch.ones(3, 64)
    cs = torch.zeros(64, d)
    row = torch.zeros(64, dtype=torch.int64)
    pos = torch.tensor([0, 5, 6], dtype=torch.int32)
    indptr = torch.tensor([0, 1, 3, 5], dtype=torch.int32)
    indices = torch.tensor([4, 2, 0, 1, 3], dtype=torch.int32)
    dl.decode_prep(torch.tensor([1, 2, 9]), embed, res, ss, pos, indptr, indices, page, 100, 10, cs, row, d, rope_theta=1e4)
    assert torch.equal(res, embed[[1, 2, 9]])
    torch.testing.assert_close(ss[0, :m], embed[[1, 2, 9]].pow(2).sum(-1))
    assert float(ss[1:].abs().sum()) == 0.0
    assert row[:m].tolist() == [4 * 100 + 0, 0 * 100 + 10, 3 * 100 + 20]
    torch.testing.assert_close(cs[1, 0], torch.cos(torch.tensor(5.0)))
    torch.testing.assert_close(cs[1, d // 2], torch.sin(torch.tensor(5.0)))
