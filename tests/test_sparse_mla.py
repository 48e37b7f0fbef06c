"""Sparse (top-k) MLA decode: trtllm_batch_decode_with_kv_cache_mla(sparse_mla_top_k=K) against a dense torch oracle that attends
only the selected rows (port of reference tests/attention/test_trtllm_gen_mla.py::test_trtllm_batch_decode_mla_sparse)."""
import math

import pytest
import torch

import flashinfer_b200 as fi


def _case(device, dtype, batch, q_len, heads, top_k, seq, page=32, holes=False):
    torch.manual_seed(0)
    n_pages = batch * ((seq + page - 1) // page) + 3
    kv = (torch.randn(n_pages, page, 576, device=device).clamp(-1, 1)).to(dtype)
    q = (torch.randn(batch, q_len, heads, 576, device=device) * 0.5).to(dtype)
    perm = torch.randperm(n_pages)
    idx = torch.full((batch, q_len, top_k), -1, dtype=torch.int32)
    for b in range(batch):
        pages = perm[b * ((seq + page - 1) // page):(b + 1) * ((seq + page - 1) // page)]
        for j in range(q_len):
            n_sel = min(top_k, seq - (q_len - 1 - j))  # later query tokens see more of the sequence
            pos = torch.randperm(seq - (q_len - 1 - j))[:n_sel]
            rows = pages[pos // page] * page + pos % page
            if holes and n_sel > 4:  # unused (-1) slots in the middle of the list
                slots = torch.randperm(top_k)[:n_sel].sort().values
                idx[b, j, slots] = rows.int()
            else:
                idx[b, j, :n_sel] = rows.int()
    return q, kv, idx.to(device)


def _oracle(q, kv, idx, scale):
    b, ql, h, _ = q.shape
    flat = kv.reshape(-1, 576).float()
    out = torch.zeros(b, ql, h, 512, device=q.device)
    for i in range(b):
        for j in range(ql):
            sel = idx[i, j][idx[i, j] >= 0].long()
            rows = flat[sel]
            logits = q[i, j].float() @ rows.t() * scale
            out[i, j] = torch.softmax(logits, -1) @ rows[:, :512]
    return out


@pytest.mark.parametrize("holes", [False, True])
def test_sparse_mla_cpu(holes):
    q, kv, idx = _case("cpu", torch.float32, 2, 2, 4, 48, 100, holes=holes)
    scale = 1 / math.sqrt(128 + 64)
    ws = torch.zeros(1 << 20, dtype=torch.uint8)
    out = fi.mla.trtllm_batch_decode_with_kv_cache_mla(q, kv.unsqueeze(1), ws, 128, 512, 64, idx, torch.full((2,), 100), 100,
                                                       sparse_mla_top_k=48, bmm1_scale=scale)
    assert (out - _oracle(q, kv, idx, scale)).abs().max() < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("top_k,holes", [(128, False), (2048, False), (200, True)])
def test_sparse_mla_gpu(top_k, holes):
    q, kv, idx = _case("cuda", torch.bfloat16, 3, 2, 128, top_k, 3000, page=64, holes=holes)
    scale = 1 / math.sqrt(128 + 64)
    ws = torch.zeros(128 << 20, dtype=torch.uint8, device="cuda")
    out, lse = fi.mla.trtllm_batch_decode_with_kv_cache_mla(q, kv.unsqueeze(1), ws, 128, 512, 64, idx,
                                                            torch.full((3,), 3000, device="cuda"), 3000, sparse_mla_top_k=top_k,
                                                            bmm1_scale=scale, return_lse=True)
    ref = _oracle(q, kv, idx, scale)
    assert out.shape == (3, 2, 128, 512) and not torch.isnan(out).any()
    assert (out.float() - ref).abs().max() < 2e-2
    with pytest.raises(ValueError):
        fi.mla.trtllm_batch_decode_with_kv_cache_mla(q, kv.unsqueeze(1), ws, 128, 512, 64, idx[:, :1], torch.full((3,), 3000), 3000,
                                                     sparse_mla_top_k=top_k, bmm1_scale=scale)
