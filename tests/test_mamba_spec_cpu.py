"""Speculative-decoding forms of mamba.selective_state_update (intermediate state caching, varlen, num_accepted_tokens, 2-D state
indices; reference tests/mamba/test_selective_state_update_{mtp,varlen}.py) against a token-by-token loop."""
import pytest
import torch

from flashinfer_b200.mamba import selective_state_update

H, DIM, DS, G, POOL = 4, 8, 16, 2, 24
PAD = -1


def _weights(seed=0):
    g = torch.Generator().manual_seed(seed)
    return dict(A=-torch.rand(H, DIM, DS, generator=g), D=torch.randn(H, DIM, generator=g), dt_bias=torch.randn(H, DIM, generator=g) * 0.1)


def _tokens(total, seed=1):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    return dict(x=r(total, H, DIM), dt=r(total, H, DIM) * 0.3, B=r(total, G, DS), C=r(total, G, DS), z=r(total, H, DIM))


def _step(s, w, tk, t):
    """One token on an fp32 state [H, DIM, DS]; returns (new state, y)."""
    dt = torch.nn.functional.softplus(tk["dt"][t] + w["dt_bias"])
    Bv, Cv = tk["B"][t].repeat_interleave(H // G, 0), tk["C"][t].repeat_interleave(H // G, 0)
    s = s * torch.exp(w["A"] * dt[..., None]) + (dt * tk["x"][t])[..., None] * Bv[:, None, :]
    y = (s * Cv[:, None, :]).sum(-1) + w["D"] * tk["x"][t]
    return s, y * torch.nn.functional.silu(tk["z"][t])


def test_intermediate_states_buffer_multi_token():
    w, bsz, T = _weights(), 3, 4
    tk = _tokens(bsz * T)
    state = (torch.randn(POOL, H, DIM, DS) * 0.1).bfloat16()
    idx = torch.tensor([7, PAD, 2], dtype=torch.int32)
    rows = torch.tensor([2, 0, 1])
    buf = torch.zeros(3, T + 1, H, DIM, DS)
    shape = lambda t: t.view(bsz, T, *t.shape[1:])  # noqa: E731
    for disable in (True, False):
        st = state.clone()
        y = selective_state_update(st, shape(tk["x"]), shape(tk["dt"]), w["A"], shape(tk["B"]), shape(tk["C"]), w["D"], shape(tk["z"]), w["dt_bias"], True,
                                   state_batch_indices=idx, pad_slot_id=PAD, disable_state_update=disable, intermediate_states_buffer=buf,
                                   intermediate_state_indices=rows, cache_steps=T + 1)
        assert y.shape == (bsz, T, H, DIM)
        for b in (0, 2):
            s = state[int(idx[b])].float()
            for t in range(T):
                s, want = _step(s, w, tk, b * T + t)
                torch.testing.assert_close(y[b, t], want, atol=1e-5, rtol=1e-5)
                torch.testing.assert_close(buf[int(rows[b]), t], s, atol=1e-5, rtol=1e-5)
            if disable:
                assert torch.equal(st, state)
            else:
                torch.testing.assert_close(st[int(idx[b])], s.bfloat16())
        assert torch.count_nonzero(y[1]) == 0 and torch.count_nonzero(buf[0]) == 0        # the padded request
    with pytest.raises(ValueError):
        selective_state_update(state.clone(), shape(tk["x"]), shape(tk["dt"]), w["A"], shape(tk["B"]), shape(tk["C"]), w["D"], None, w["dt_bias"], True,
                               intermediate_states_buffer=torch.zeros(3, T - 1, H, DIM, DS))


@pytest.mark.parametrize("lens", [[3, 3, 3], [1, 4, 2, 0, 3]])
def test_varlen_with_num_accepted_tokens(lens):
    """Sequence n starts from the slot of its last accepted token and leaves the state after token t in dst[n, t]."""
    w = _weights(2)
    n, mx = len(lens), max(lens)
    cu = torch.tensor([0] + torch.tensor(lens).cumsum(0).tolist(), dtype=torch.int32)
    tk = _tokens(int(cu[-1]), seed=3)
    state = (torch.randn(POOL, H, DIM, DS) * 0.1).bfloat16()
    perm = torch.randperm(POOL, generator=torch.Generator().manual_seed(4))
    src = perm[: n * mx].view(n, mx).int()
    dst = perm[n * mx: 2 * n * mx].view(n, mx).int() if 2 * n * mx <= POOL else src.clone()
    dst[0, 0] = PAD                                                                     # one store suppressed
    acc = torch.tensor([1 + (i % mx) for i in range(n)])
    st = state.clone()
    y = selective_state_update(st, tk["x"], tk["dt"], w["A"], tk["B"], tk["C"], w["D"], tk["z"], w["dt_bias"], True, state_batch_indices=src,
                               dst_state_batch_indices=dst, pad_slot_id=PAD, cu_seqlens=cu, num_accepted_tokens=acc, cache_steps=mx)
    want_state = state.clone()
    for i in range(n):
        s = state[int(src[i, int(acc[i]) - 1])].float()
        for t in range(lens[i]):
            s, wy = _step(s, w, tk, int(cu[i]) + t)
            torch.testing.assert_close(y[int(cu[i]) + t], wy, atol=1e-5, rtol=1e-5)
            if int(dst[i, t]) != PAD:
                want_state[int(dst[i, t])] = s.bfloat16()
    torch.testing.assert_close(st, want_state)


def test_varlen_without_spec_decoding_writes_the_final_state():
    w = _weights(5)
    lens = [2, 5, 1]
    cu = torch.tensor([0, 2, 7, 8], dtype=torch.int32)
    tk = _tokens(8, seed=6)
    state = torch.randn(POOL, H, DIM, DS) * 0.1
    src, dst = torch.tensor([3, 9, 11], dtype=torch.int32), torch.tensor([4, 9, 0], dtype=torch.int32)
    st = state.clone()
    out = torch.empty(8, H, DIM)
    ret = selective_state_update(st, tk["x"], tk["dt"], w["A"], tk["B"], tk["C"], w["D"], tk["z"], w["dt_bias"], True, state_batch_indices=src,
                                 dst_state_batch_indices=dst, cu_seqlens=cu, out=out)
    assert ret is out
    want_state = state.clone()
    for i in range(3):
        s = state[int(src[i])]
        for t in range(lens[i]):
            s, wy = _step(s, w, tk, int(cu[i]) + t)
            torch.testing.assert_close(out[int(cu[i]) + t], wy, atol=1e-5, rtol=1e-5)
        want_state[int(dst[i])] = s
    torch.testing.assert_close(st, want_state, atol=1e-5, rtol=1e-5)
    # equal-length sequences in the varlen form == the multi-token form
    cu2 = torch.tensor([0, 4, 8], dtype=torch.int32)
    s1, s2 = state.clone(), state.clone()
    y1 = selective_state_update(s1, tk["x"], tk["dt"], w["A"], tk["B"], tk["C"], w["D"], tk["z"], w["dt_bias"], True, state_batch_indices=src[:2], cu_seqlens=cu2)
    v4 = lambda t: t.view(2, 4, *t.shape[1:])  # noqa: E731
    y2 = selective_state_update(s2, v4(tk["x"]), v4(tk["dt"]), w["A"], v4(tk["B"]), v4(tk["C"]), w["D"], v4(tk["z"]), w["dt_bias"], True, state_batch_indices=src[:2])
    torch.testing.assert_close(y1.view(2, 4, H, DIM), y2, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(s1, s2, atol=1e-5, rtol=1e-5)
    with pytest.raises(NotImplementedError):
        selective_state_update(s1, tk["x"], tk["dt"], w["A"], tk["B"], tk["C"], w["D"], state_scale=torch.ones(POOL, H, DIM))
