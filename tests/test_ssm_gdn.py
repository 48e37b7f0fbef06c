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
        out, _ = gated_delta_rule_mtp(q, k, v, pool, idx, A_log, a, dt_bias, b, disable_state_update=False, state_layout="KV")
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
    out, state = chunk_gated_delta_rule(q, k, v, g, beta, cu_seqlens=cu, output_final_state=True, use_qk_l2norm_in_kernel=True, state_layout="KV")
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
    init = torch.randn(3, HV, V, K) * 0.1                     # K-last, the reference's state layout
    o1, s1 = chunk_gated_delta_rule(q, k, v, g, beta, initial_state=init.clone(), output_final_state=True, cu_seqlens=cu,
                                    use_qk_l2norm_in_kernel=True)
    o2, s2 = chunk_gated_delta_rule(q, k, v, g, beta, initial_state=init.clone(), output_final_state=True, cu_seqlens=cu,
                                    use_qk_l2norm_in_kernel=True, chunked=True, chunk_size=32)
    torch.testing.assert_close(o2.float(), o1.float(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(s2.float(), s1.float(), rtol=1e-4, atol=1e-5)


def _gdn_inputs(B, T, H, HV, K, V, seed=0):
    import torch

    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    kk = r(B, T, H, K)
    return dict(q=r(B, T, H, K), k=kk / kk.norm(dim=-1, keepdim=True), v=r(B, T, HV, V), a=r(B, T, HV), b=r(B, T, HV), A_log=r(HV) * 0.5, dt_bias=r(HV) * 0.1)


def test_gdn_mtp_state_conventions_cpu():
    """Reference conventions of gated_delta_rule_mtp: K-last pool, state NOT updated by default (with a warning), per-token
    intermediate states indexed by the row of the call; the native K-major layout gives the same numbers."""
    import warnings

    import torch
    from flashinfer_b200.gdn import gated_delta_rule_decode, gated_delta_rule_mtp, gated_delta_rule_ref

    B, T, H, HV, K, V = 3, 4, 2, 4, 16, 8
    x = _gdn_inputs(B, T, H, HV, K, V)
    pool_kv = torch.randn(6, HV, K, V) * 0.1
    pool_vk = pool_kv.transpose(-1, -2).contiguous()
    idx = torch.tensor([4, 0, 5], dtype=torch.int32)
    g_log = -torch.exp(x["A_log"]) * torch.nn.functional.softplus(x["a"] + x["dt_bias"])
    st = pool_kv[idx.long()].clone()
    want = gated_delta_rule_ref(x["q"], x["k"], x["v"], st, g_log, torch.sigmoid(x["b"]), K ** -0.5, True)
    args = (x["q"], x["k"], x["v"])
    tail = (x["A_log"], x["a"], x["dt_bias"], x["b"])
    before = pool_vk.clone()
    with pytest.warns(FutureWarning):
        o, _ = gated_delta_rule_mtp(*args, pool_vk, idx, *tail)
    torch.testing.assert_close(o, want, atol=1e-5, rtol=1e-5)
    assert torch.equal(pool_vk, before)                                  # default: the pool is left alone
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        buf = torch.zeros(B, T + 1, HV, V, K)
        o, ret = gated_delta_rule_mtp(*args, pool_vk, idx, *tail, intermediate_states_buffer=buf, disable_state_update=False)
    assert ret is pool_vk
    torch.testing.assert_close(o, want, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(pool_vk[idx.long()], st.transpose(-1, -2), atol=1e-5, rtol=1e-5)
    untouched = [i for i in range(6) if i not in idx.tolist()]
    assert torch.equal(pool_vk[untouched], before[untouched])
    torch.testing.assert_close(buf[:, T - 1], st.transpose(-1, -2), atol=1e-5, rtol=1e-5)
    assert torch.count_nonzero(buf[:, T]) == 0
    # buffer row t is the state a single-token decode chain reaches after t + 1 tokens
    chain = pool_kv[idx.long()].clone()
    for t in range(2):
        gated_delta_rule_decode(x["q"][:, t:t + 1], x["k"][:, t:t + 1], x["v"][:, t:t + 1], chain, x["A_log"], x["a"][:, t:t + 1], x["dt_bias"], x["b"][:, t:t + 1])
    torch.testing.assert_close(buf[:, 1], chain.transpose(-1, -2), atol=1e-5, rtol=1e-5)
    # native layout
    pk = pool_kv.clone()
    o2, _ = gated_delta_rule_mtp(*args, pk, idx, *tail, disable_state_update=False, state_layout="KV")
    torch.testing.assert_close(o2, want, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(pk[idx.long()], st, atol=1e-5, rtol=1e-5)
    buf2 = torch.zeros(B, T, HV, V, K)
    pk2 = pool_kv.clone()
    gated_delta_rule_mtp(*args, pk2, idx, *tail, disable_state_update=True, state_layout="KV", intermediate_states_buffer=buf2)
    assert torch.equal(pk2, pool_kv)
    torch.testing.assert_close(buf2, buf[:, :T], atol=1e-5, rtol=1e-5)
    with pytest.raises(ValueError):
        gated_delta_rule_mtp(*args, pk2, idx, *tail, disable_state_update=True, state_layout="NK")
    with pytest.raises(ValueError):
        gated_delta_rule_mtp(*args, pool_vk, idx, *tail, disable_state_update=True, intermediate_states_buffer=torch.zeros(B, T - 1, HV, V, K))


def test_gdn_decode_pretranspose_pool_form_cpu():
    import torch
    from flashinfer_b200.gdn import gated_delta_rule_decode, gated_delta_rule_decode_pretranspose

    B, H, HV, K, V = 4, 2, 4, 16, 8
    x = _gdn_inputs(B, 1, H, HV, K, V, seed=3)
    pool = torch.randn(7, HV, V, K) * 0.1
    before = pool.clone()
    src = torch.tensor([2, -1, 6, 1], dtype=torch.int32)
    dst = torch.tensor([3, -1, 6, 0], dtype=torch.int64)
    st = pool[src.clamp(min=0).long()].transpose(-1, -2).contiguous()
    want, _ = gated_delta_rule_decode(x["q"], x["k"], x["v"], st, x["A_log"], x["a"], x["dt_bias"], x["b"])
    o, ret = gated_delta_rule_decode_pretranspose(x["q"], x["k"], x["v"], None, x["A_log"], x["a"], x["dt_bias"], x["b"], initial_state=pool,
                                                  initial_state_indices=src, output_state_indices=dst)
    assert ret is pool
    live = [0, 2, 3]
    torch.testing.assert_close(o[live], want[live])
    assert torch.count_nonzero(o[1]) == 0
    torch.testing.assert_close(pool[dst[live]], st[live].transpose(-1, -2))
    rest = [i for i in range(7) if i not in dst[live].tolist()]
    assert torch.equal(pool[rest], before[rest])                          # incl. the read-only source slots 2 and 1
    # per-batch state form still works and agrees
    sv = before[src.clamp(min=0).long()].clone()
    o2, _ = gated_delta_rule_decode_pretranspose(x["q"], x["k"], x["v"], sv, x["A_log"], x["a"], x["dt_bias"], x["b"])
    torch.testing.assert_close(o2[live], want[live])
    with pytest.raises(ValueError):
        gated_delta_rule_decode_pretranspose(x["q"], x["k"], x["v"], None, x["A_log"], x["a"], x["dt_bias"], x["b"])
    with pytest.raises(ValueError):
        gated_delta_rule_decode_pretranspose(x["q"], x["k"], x["v"], sv, x["A_log"], x["a"], x["dt_bias"], x["b"], initial_state=pool, initial_state_indices=src)


@pytest.mark.parametrize("chunked", [False, True])
def test_gdn_prefill_layouts_and_checkpoints_cpu(chunked):
    """K-last states by default, native layout on request; state checkpoints every n tokens equal the final states of the prefixes."""
    import torch
    from flashinfer_b200.gdn import chunk_gated_delta_rule

    torch.manual_seed(4)
    H, HV, K, V = 2, 4, 16, 8
    lens = [150, 64, 30, 130]
    cu = torch.tensor([0] + torch.tensor(lens).cumsum(0).tolist(), dtype=torch.int32)
    total = cu[-1].item()
    q, k, v = torch.randn(total, H, K), torch.randn(total, H, K), torch.randn(total, HV, V)
    g, beta = torch.exp(-torch.rand(total, HV) * 0.3), torch.rand(total, HV)
    init_vk = torch.randn(4, HV, V, K) * 0.1
    kw = dict(cu_seqlens=cu, use_qk_l2norm_in_kernel=True, output_final_state=True, chunked=chunked)
    o_vk, s_vk = chunk_gated_delta_rule(q, k, v, g, beta, initial_state=init_vk, **kw)
    o_kv, s_kv = chunk_gated_delta_rule(q, k, v, g, beta, initial_state=init_vk.transpose(-1, -2).contiguous(), state_layout="KV", **kw)
    assert s_vk.shape == (4, HV, V, K) and s_kv.shape == (4, HV, K, V) and s_vk.is_contiguous()
    torch.testing.assert_close(o_vk, o_kv)
    torch.testing.assert_close(s_vk, s_kv.transpose(-1, -2))
    buf = torch.full((4, HV, V, K), 7.0)
    _, ret = chunk_gated_delta_rule(q, k, v, g, beta, initial_state=init_vk, output_state=buf, **kw)
    assert ret is buf
    torch.testing.assert_close(buf, s_vk)
    nbuf = torch.full((4, HV, K, V), 7.0)
    chunk_gated_delta_rule(q, k, v, g, beta, output_state=nbuf, state_layout="KV", **kw)       # no initial state: starts from zero
    _, s0 = chunk_gated_delta_rule(q, k, v, g, beta, **kw)
    torch.testing.assert_close(nbuf, s0.transpose(-1, -2))
    # checkpoints every 64 tokens: 2 + 1 + 0 + 2
    starts = torch.tensor([0, 2, 3, 3, 5], dtype=torch.int64)
    ck = torch.full((5, HV, V, K), float("nan"))
    o_ck, s_ck = chunk_gated_delta_rule(q, k, v, g, beta, initial_state=init_vk, state_checkpoints=ck, checkpoint_cu_starts=starts,
                                        checkpoint_every_n_tokens=64, **kw)
    torch.testing.assert_close(o_ck, o_vk, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(s_ck, s_vk, atol=1e-5, rtol=1e-5)
    for i, row, n_tok in [(0, 0, 64), (0, 1, 128), (1, 2, 64), (3, 3, 64), (3, 4, 128)]:
        sl = slice(int(cu[i]), int(cu[i]) + n_tok)
        _, pre = chunk_gated_delta_rule(q[sl], k[sl], v[sl], g[sl], beta[sl], initial_state=init_vk[i:i + 1], use_qk_l2norm_in_kernel=True,
                                        output_final_state=True, chunked=chunked)
        torch.testing.assert_close(ck[row], pre[0], atol=1e-5, rtol=1e-5)
    with pytest.raises(ValueError):
        chunk_gated_delta_rule(q, k, v, g, beta, state_checkpoints=ck, checkpoint_cu_starts=starts, checkpoint_every_n_tokens=100, **kw)
    with pytest.raises(ValueError):
        chunk_gated_delta_rule(q, k, v, g, beta, checkpoint_every_n_tokens=64, **kw)
    with pytest.raises(ValueError):
        chunk_gated_delta_rule(q, k, v, g, beta, state_checkpoints=ck, checkpoint_cu_starts=torch.tensor([0, 1, 2, 3, 5]), checkpoint_every_n_tokens=64, **kw)
