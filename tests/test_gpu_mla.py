"""GPU tests of the tcgen05 MLA decode kernel vs. the fp32 oracle."""
import pytest
import torch

from flashinfer_b200.mla import BatchMLAPagedAttentionWrapper, mla_attention_ref, trtllm_batch_decode_with_kv_cache_mla

pytestmark = pytest.mark.gpu


def _run(kv_lens, q_lens, H, ps, dtype, causal=True):
    B = len(kv_lens)
    npg = [(l + ps - 1) // ps for l in kv_lens]
    kvp = torch.tensor([0] + torch.tensor(npg).cumsum(0).tolist(), dtype=torch.int32)
    idx = torch.randperm(sum(npg) + 2)[: sum(npg)].int()
    ckv = (torch.randn(sum(npg) + 2, ps, 512, device="cuda") * 0.5).to(dtype)
    kpe = (torch.randn(sum(npg) + 2, ps, 64, device="cuda") * 0.5).to(dtype)
    qo = torch.tensor([0] + torch.tensor(q_lens).cumsum(0).tolist(), dtype=torch.int32)
    n = int(qo[-1])
    qn = (torch.randn(n, H, 512, device="cuda") * 0.5).to(dtype)
    qp = (torch.randn(n, H, 64, device="cuda") * 0.5).to(dtype)
    sm_scale = 1.0 / (192 ** 0.5)
    w = BatchMLAPagedAttentionWrapper(torch.empty(128 << 20, dtype=torch.uint8, device="cuda"))
    w.plan(qo, kvp, idx, torch.tensor(kv_lens, dtype=torch.int32), H, 512, 64, ps, causal, sm_scale, dtype, dtype)
    o, lse = w.run(qn, qp, ckv, kpe, return_lse=True)
    for b in range(B):
        qs, qe = int(qo[b]), int(qo[b + 1])
        pages = idx[int(kvp[b]) : int(kvp[b + 1])].long().cuda()
        c = ckv[pages].reshape(-1, 512)[: kv_lens[b]]
        k = kpe[pages].reshape(-1, 64)[: kv_lens[b]]
        o_ref, l_ref = mla_attention_ref(qn[qs:qe], qp[qs:qe], c, k, sm_scale, kv_lens[b] - (qe - qs) if causal else None)
        torch.testing.assert_close(o[qs:qe].float(), o_ref, rtol=3e-2, atol=3e-2)
        torch.testing.assert_close(lse[qs:qe], l_ref, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("cfg", [
    ([70], [1], 128, 32, torch.bfloat16),
    ([70, 200, 33], [1, 1, 1], 16, 32, torch.bfloat16),
    ([1024] * 4, [1] * 4, 128, 64, torch.float16),
    ([300, 5000, 17], [1, 1, 1], 128, 16, torch.bfloat16),
    ([129, 64], [2, 3], 64, 32, torch.bfloat16),
    ([4096] * 2, [1, 1], 128, 128, torch.bfloat16),
    ([100], [1], 128, 8, torch.bfloat16),
], ids=lambda c: f"kv{c[0][:2]}x{len(c[0])}-q{c[1][:2]}-h{c[2]}-ps{c[3]}")
def test_mla_decode(cfg):
    _run(*cfg)


def test_trtllm_mla_function_api():
    B, ql, H, ps = 3, 1, 128, 32
    seq = torch.tensor([100, 333, 32], dtype=torch.int32, device="cuda")
    max_pages = 12
    kv = (torch.randn(B * max_pages, ps, 576, device="cuda") * 0.5).bfloat16()
    bt = torch.arange(B * max_pages, device="cuda", dtype=torch.int32).view(B, max_pages)
    q = (torch.randn(B, ql, H, 576, device="cuda") * 0.5).bfloat16()
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    o = trtllm_batch_decode_with_kv_cache_mla(q, kv, ws, 128, 512, 64, bt, seq, 333, bmm1_scale=0.07)
    for b in range(B):
        n = int(seq[b])
        flat = kv[bt[b].long()].reshape(-1, 576)[:n]
        o_ref, _ = mla_attention_ref(q[b, :, :, :512], q[b, :, :, 512:], flat[:, :512], flat[:, 512:], 0.07, n - ql)
        torch.testing.assert_close(o[b].float(), o_ref, rtol=3e-2, atol=3e-2)
