"""Block-sparse attention wrappers against dense masked attention (port of reference tests/attention/test_block_sparse.py):
BSR masks with (R, C) blocks incl. R > 1 with small GQA groups, block widths that are not powers of two, element-level masks
inside the blocks (dense and bit-packed), and the variable-block-size wrapper."""
import math

import pytest
import torch

import flashinfer_b200 as fi
from flashinfer_b200 import reference


def _bsr(mb, nb, density, seed):
    g = torch.Generator().manual_seed(seed)
    dense = torch.rand(mb, nb, generator=g) < density
    dense[:, 0] = True  # every row attends at least one block
    indptr = torch.zeros(mb + 1, dtype=torch.int32)
    indptr[1:] = dense.sum(1).cumsum(0)
    indices = dense.nonzero()[:, 1].int()
    return dense, indptr, indices


def _run_case(device, dtype, R, C, hq, hkv, d, mb, nb, elem_mask=False, packed=False, tol=2e-2):
    torch.manual_seed(R * 100 + C)
    M, N = mb * R, nb * C
    dense, indptr, indices = _bsr(mb, nb, 0.4, seed=R + C)
    q = torch.randn(M, hq, d, device=device).to(dtype)
    k = torch.randn(N, hkv, d, device=device).to(dtype)
    v = torch.randn(N, hkv, d, device=device).to(dtype)
    full = dense.repeat_interleave(R, 0).repeat_interleave(C, 1).to(device)
    kw = {}
    if elem_mask:
        em = torch.rand(int(indices.numel()), R, C) > 0.3
        em[:, :, 0] = True
        # scatter the block masks into the dense oracle mask
        blk = 0
        for i in range(mb):
            for j in indices[int(indptr[i]):int(indptr[i + 1])].tolist():
                full[i * R:(i + 1) * R, j * C:(j + 1) * C] = em[blk].to(device)
                blk += 1
        if packed:
            flat = fi.sparse.convert_bsr_mask_layout(em, indptr)
            seg = (indptr.long() * (R * C)).int()     # reference format: one byte-aligned segment per block row (segment_packbits)
            kw["packed_mask"] = fi.segment_packbits(flat.cpu(), seg, bitorder="little")[0].to(device)
        else:
            kw["mask"] = em.to(device)
    w = fi.BlockSparseAttentionWrapper(torch.empty(32 << 20, dtype=torch.uint8, device=device))
    w.plan(indptr, indices, M, N, R, C, hq, hkv, d, q_data_type=dtype, **kw)
    out = w.run(q, k, v)
    ref, _ = reference.attention_ref(q, k, v, False, 1 / math.sqrt(d), custom_mask=full)
    assert (out.float() - ref.float()).abs().max() < tol


@pytest.mark.parametrize("R,C,elem,packed", [(1, 16, False, False), (4, 8, False, False), (16, 16, True, False), (2, 24, False, False),
                                             (3, 3, True, True)])          # 9 bits per block: segments end inside a byte
def test_block_sparse_cpu(R, C, elem, packed):
    _run_case("cpu", torch.float32, R, C, 4, 2, 32, 6, 5, elem_mask=elem, packed=packed, tol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("R,C,hq,hkv,elem,packed", [(1, 64, 8, 2, False, False),      # decode kernel
                                                    (4, 64, 8, 8, False, False),      # R > 1, group 1: prefill kernel
                                                    (16, 16, 8, 2, True, False),      # element masks inside the blocks
                                                    (64, 32, 4, 4, True, True),       # bit-packed element masks
                                                    (128, 128, 8, 2, False, False),
                                                    (8, 48, 8, 2, False, False),      # C = 48: pages of 16 tokens
                                                    (2, 24, 4, 4, True, False)])      # C = 24: pages of 8 tokens + element masks
def test_block_sparse_gpu(R, C, hq, hkv, elem, packed):
    _run_case("cuda", torch.bfloat16, R, C, hq, hkv, 128, 9, 11, elem_mask=elem, packed=packed)


def _variable_case(device, dtype, causal, tol):
    torch.manual_seed(5)
    hq, hkv, d = 4, 2, 128 if device == "cuda" else 32
    mb, nb = 5, 6
    row_sz = torch.tensor([[3, 17, 64, 1, 40], [10, 30, 5, 70, 10]])
    col_sz = torch.tensor([[16, 1, 50, 33, 20, 8], [8, 20, 33, 50, 1, 16]])
    bm = torch.rand(hkv, mb, nb) < 0.5
    bm[:, :, 0] = True
    sq, skv = int(row_sz[0].sum()), int(col_sz[0].sum())
    q = torch.randn(hq, sq, d, device=device).to(dtype)
    k = torch.randn(hkv, skv, d, device=device).to(dtype)
    v = torch.randn(hkv, skv, d, device=device).to(dtype)
    w = fi.VariableBlockSparseAttentionWrapper(torch.empty(32 << 20, dtype=torch.uint8, device=device))
    w.plan(bm, row_sz, col_sz, hq, hkv, d, causal=causal, q_data_type=dtype)
    out = w.run(q, k, v)
    g = hq // hkv
    for h in range(hkv):
        rs, cs = row_sz[h].tolist(), col_sz[h].tolist()
        r0 = 0
        for i in range(mb):
            cols = []
            c0 = 0
            for j in range(nb):
                if bm[h, i, j]:
                    cols += list(range(c0, c0 + cs[j]))
                c0 += cs[j]
            ci = torch.tensor(cols, device=device)
            qq = q[h * g:(h + 1) * g, r0:r0 + rs[i]].transpose(0, 1)          # [rows, g, d]
            kk, vv = k[h, ci][:, None], v[h, ci][:, None]                      # [cols, 1, d]
            ref, _ = reference.attention_ref(qq, kk, vv, causal, 1 / math.sqrt(d))
            got = out[h * g:(h + 1) * g, r0:r0 + rs[i]].transpose(0, 1)
            assert (got.float() - ref.float()).abs().max() < tol
            r0 += rs[i]


@pytest.mark.parametrize("causal", [False, True])
def test_variable_block_sparse_cpu(causal):
    _variable_case("cpu", torch.float32, causal, 1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("causal", [False, True])
def test_variable_block_sparse_gpu(causal):
    _variable_case("cuda", torch.bfloat16, causal, 2e-2)
