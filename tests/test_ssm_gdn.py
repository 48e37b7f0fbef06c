"""Mamba selective_state_update, gated delta rule, concat_mla_k, router GEMMs vs fp32 PyTorch oracles
(reference tests/mamba/test_selective_state_update*.py, tests/gdn/*, tests/utils/test_concat_mla.py)."""
import pytest
import torch

from flashinfer_b200.concat_ops import concat_mla_k
from flashinfer_b200.gdn import (chunk_gated_delta_rule, gated_delta_rule_decode, gated_delta_rule_mtp, gated_delta_rule_ref)
from flashinfer_b200.mamba import selective_state_update, selective_state_update_ref


def test_ssu_cpu_shapes():
    st = torch.randn(4, 16, 32)
    x, dt = torch.randn(3, 16), torch.randn(3, 16)
    out = selective_state_update(st, x, dt, -torch.rand(16, 32), torch.randn(3, 32), torch.randn(3, 32), torch.randn(16),
                                 state_batch_indices=torch.tensor([0, 2, 3]))
    assert out.shape == (3, 16)


@pytest.mark.gpu
@pytest.mark.parametrize("dstate", [16, 64, 128])
@pytest.mark.parametrize("T", [1, 4])
@pytest.mark.parametrize("state_dtype", [torch.float32, torch.bfloat16])
def test_ssu_gpu(dstate, T, state_dtype):
    B, H, dim, G, N = 7, 8, 64, 2, 16
    torch.manual_seed(0)
    state = torch.randn(N, H, dim, dstate, device="cuda").to(state_dtype)
    x = torch.randn(B, T, H, dim, device="cuda", dtype=torch.bfloat16)
    dt = torch.randn(B, T, H, device="cuda", dtype=torch.bfloat16)[..., None].expand(B, T, H, dim)  # tie_hdim
    A = (-torch.rand(H, device="cuda"))[:, None, None].expand(H, dim, dstate)
    Bm = torch.randn(B, T, G, dstate, device="cuda", dtype=torch.bfloat16)
    Cm = torch.randn(B, T, G, dstate, device="cuda", dtype=torch.bfloat16)
    D = torch.randn(H, device="cuda")[:, None].expand(H, dim)
    z = torch.randn(B, T, H, dim, device="cuda", dtype=torch.bfloat16)
    dtb = torch.randn(H, device="cuda")[:, None].expand(H, dim)
    idx = torch.randperm(N, device="cuda")[:B].int()
    idx[2] = -1
    ref_state = state.clone()
    ref = selective_state_update_ref(ref_state, x, dt, A, Bm, Cm, D, z, dtb, True, idx, -1)
    if T == 1:
        out = selective_state_update(state, x[:, 0], dt[:, 0], A, Bm[:, 0], Cm[:, 0], D, z[:, 0], dtb, True, idx, -1)[:, None]
    else:
        out = selective_state_update(state, x, dt, A, Bm, Cm, D, z, dtb, True, idx, -1)
    keep = idx >= 0
    assert (out.float() - ref)[keep].abs().max() < 5e-2 * max(1.0, float(ref.abs().max()))
    tol = 1e-4 if state_dtype == torch.float32 else 3e-2
    assert (state.float() - ref_state.float()).abs().max() < tol * max(1.0, float(ref_state.float().abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("K,V", [(128, 128), (64, 64), (128, 64)])
@pytest.mark.parametrize("T", [1, 5])
def test_gdn_decode_gpu(K, V, T):
    B, H, HV = 5, 4, 8
    torch.manual_seed(1)
    q = torch.randn(B, T, H, K, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(B, T, H, K, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(B, T, HV, V, device="cuda", dtype=torch.bfloat16)
    a = torch.randn(B, T, HV, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(B, T, HV, device="cuda", dtype=torch.bfloat16)
    A_log = torch.randn(HV, device="cuda") * 0.5
    dt_bias = torch.randn(HV, device="cuda") * 0.5
    pool = torch.randn(9, HV, K, V, device="cuda") * 0.1
    idx = torch.tensor([3, 0, 8, 5, 1], device="cuda", dtype=torch.int32)
    g_log = -torch.exp(A_log) * torch.nn.functional.softplus(a.float() + dt_bias)
    ref_state = pool[idx.long()].clone()
    ref = gated_delta_rule_ref(q, k, v, ref_state, g_log, torch.sigmoid(b.float()), K ** -0.5, True)
    if T == 1:
        st = pool[idx.long()].clone()
        out, st = gated_delta_rule_decode(q, k, v, st, A_log, a, dt_bias, b)
        got_state = st
    else:
        out, _ = gated_delta_rule_mtp(q, k, v, pool, idx, A_log, a, dt_bias, b)
        got_state = pool[idx.long()]
    assert (out.float() - ref).abs().max() < 3e-2 * max(1.0, float(ref.abs().max()))
    assert (got_state - ref_state).abs().max() < 2e-3 * max(1.0, float(ref_state.abs().max()))


@pytest.mark.gpu
def test_gdn_prefill_varlen_gpu():
    H, HV, K, V = 4, 4, 128, 128
    cu = torch.tensor([0, 37, 100, 101], dtype=torch.int32, device="cuda")
    total = 101
    torch.manual_seed(2)
    q = torch.randn(total, H, K, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(total, H, K, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(total, HV, V, device="cuda", dtype=torch.bfloat16)
    g = torch.rand(total, HV, device="cuda") * 0.5 + 0.5
    beta = torch.rand(total, HV, device="cuda")
    out, state = chunk_gated_delta_rule(q, k, v, g, beta, cu_seqlens=cu, output_final_state=True, use_qk_l2norm_in_kernel=True)
    for i in range(3):
        sl = slice(int(cu[i]), int(cu[i + 1]))
        st = torch.zeros(1, HV, K, V, device="cuda")
        ref = gated_delta_rule_ref(q[None, sl], k[None, sl], v[None, sl], st, torch.log(g[None, sl]), beta[None, sl], K ** -0.5, True)
        assert (out[sl].float() - ref[0]).abs().max() < 3e-2 * max(1.0, float(ref.abs().max()))
        assert (state[i] - st[0]).abs().max() < 2e-3 * max(1.0, float(st.abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float8_e4m3fn])
def test_concat_mla_k_gpu(dtype):
    T, H, nope, rope = 333, 128, 128, 64
    kn = torch.randn(T, H, nope, device="cuda").to(dtype)
    kr = torch.randn(T, 1, rope, device="cuda").to(dtype)
    k = torch.empty(T, H, nope + rope, device="cuda", dtype=dtype)
    concat_mla_k(k, kn, kr)
    assert torch.equal(k[..., :nope].view(torch.uint8), kn.view(torch.uint8))
    assert torch.equal(k[..., nope:].view(torch.uint8), kr.expand(T, H, rope).contiguous().view(torch.uint8))


@pytest.mark.gpu
def test_router_gemm_gpu():
    from flashinfer_b200.dsv3_ops import mm_M1_16_K7168_N256, tinygemm_bf16

    a = (torch.randn(16, 7168, device="cuda") * 0.1).bfloat16()
    w = (torch.randn(256, 7168, device="cuda") * 0.1).bfloat16()
    out = mm_M1_16_K7168_N256(a, w.t())
    ref = a.float() @ w.float().t()
    assert (out.float() - ref).abs().max() < 2e-2 * float(ref.abs().max())
    out2 = tinygemm_bf16(a[:3], w, bias=torch.ones(256, device="cuda", dtype=torch.bfloat16))
    assert (out2.float() - (ref[:3] + 1)).abs().max() < 2e-2 * float(ref.abs().max())


def test_ssd_combined_matches_recurrence_cpu():
    """Mamba-2 chunked SSD vs the token-by-token recurrence (D, z, dt_bias + softplus, initial state, groups, packed seqs)."""
    import torch
    from flashinfer_b200.mamba import SSDCombined, ssd_reference

    torch.manual_seed(0)
    Bsz, L, H, P, G, N, Lc = 2, 64, 4, 8, 2, 16, 16
    x = torch.randn(Bsz, L, H, P)
    dt = torch.randn(Bsz, L, H) * 0.5
    A = -torch.rand(H) - 0.1
    Bm, Cm = torch.randn(Bsz, L, G, N) * 0.5, torch.randn(Bsz, L, G, N) * 0.5
    D, z, bias = torch.randn(H), torch.randn(Bsz, L, H, P), torch.randn(H) * 0.1
    init = torch.randn(Bsz, H, P, N) * 0.3
    ssd = SSDCombined(Lc, H, P, N, G, io_dtype=torch.float32, state_dtype=torch.float32)
    y, fin = ssd.run(x, dt, A, Bm, Cm, D=D, z=z, dt_bias=bias, dt_softplus=True, initial_states=init)
    y_ref, s_ref = ssd_reference(x, dt, A, Bm, Cm, D, z, bias, True, init)
    assert y.shape == (Bsz, H, P, L // Lc, Lc)
    torch.testing.assert_close(y.permute(0, 3, 4, 1, 2).reshape(Bsz, L, H, P), y_ref, rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(fin, s_ref, rtol=2e-4, atol=2e-4)
    # packed sequences: two sequences of 40 and 24 tokens in one row; the state restarts at the boundary
    seq_idx = torch.cat([torch.zeros(40), torch.ones(24)]).long()[None]
    y2, fin2 = ssd.run(x[:1], dt[:1], A, Bm[:1], Cm[:1], D=D, dt_bias=bias, dt_softplus=True, seq_idx=seq_idx)
    ya, sa = ssd_reference(x[:1, :40], dt[:1, :40], A, Bm[:1, :40], Cm[:1, :40], D, None, bias, True)
    yb, sb = ssd_reference(x[:1, 40:], dt[:1, 40:], A, Bm[:1, 40:], Cm[:1, 40:], D, None, bias, True)
    got = y2.permute(0, 3, 4, 1, 2).reshape(1, L, H, P)
    torch.testing.assert_close(got, torch.cat([ya, yb], 1), rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(fin2, torch.cat([sa, sb], 0), rtol=2e-4, atol=2e-4)


def test_chunked_gated_delta_rule_matches_sequential_cpu():
    """Chunk-parallel WY form of the gated delta rule vs the token-sequential path (packed sequences, GQA, l2norm)."""
    import torch
    from flashinfer_b200.gdn import chunk_gated_delta_rule

    torch.manual_seed(0)
    total, H, K, HV, V = 230, 2, 16, 4, 8
    q, k = torch.randn(total, H, K), torch.randn(total, H, K)
    v = torch.randn(total, HV, V)
    g = torch.exp(-torch.rand(total, HV) * 0.3)
    beta = torch.rand(total, HV)
    cu = torch.tensor([0, 100, 101, 230], dtype=torch.int32)
    init = torch.randn(3, HV, K, V) * 0.1
    o1, s1 = chunk_gated_delta_rule(q, k, v, g, beta, initial_state=init.clone(), output_final_state=True, cu_seqlens=cu,
                                    use_qk_l2norm_in_kernel=True)
    o2, s2 = chunk_gated_delta_rule(q, k, v, g, beta, initial_state=init.clone(), output_final_state=True, cu_seqlens=cu,
                                    use_qk_l2norm_in_kernel=True, chunked=True, chunk_size=32)
    torch.testing.assert_close(o2.float(), o1.float(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(s2.float(), s1.float(), rtol=1e-4, atol=1e-5)
