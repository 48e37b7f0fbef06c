"""The legacy pointer-table calling convention of the TRT-LLM custom all-reduce (reference trtllm_ar.py :430 / :809) - host-side logic
with a stand-in communicator (the real one needs GPUs; its kernels are covered by tests/test_comm_multigpu.py)."""
import pytest
import torch

from flashinfer_b200.comm import trtllm_ar as T


class _Comm:
    """World of 2 identical ranks: the sum is 2 * x."""

    def allreduce_add_rmsnorm(self, x, residual, weight, eps=1e-6, out=None, two_shot=None, **kw):
        s = 2.0 * x.float()
        if weight is None:
            out.copy_(s.to(out.dtype))
            return out
        residual.copy_((residual.float() + s).to(residual.dtype))
        r = residual.float()
        out.copy_((r * torch.rsqrt(r.pow(2).mean(-1, keepdim=True) + eps) * weight.float()).to(out.dtype))
        return out


class _Ws(T.AllReduceFusionWorkspace):
    def __init__(self, world_size, rank, max_token_num, hidden_dim, dtype, group):
        self.world_size, self.rank, self.max_token_num, self.hidden_dim, self.dtype = world_size, rank, max_token_num, hidden_dim, dtype
        self.comm, self.destroyed = _Comm(), False

    def destroy(self):
        self.destroyed = True


def test_legacy_custom_all_reduce(monkeypatch):
    monkeypatch.setattr(T, "TRTLLMAllReduceFusionWorkspace", _Ws)
    handles = T.trtllm_create_ipc_workspace_for_all_reduce(1, 2, 64, 32)
    assert len(handles) == 7 and all(len(row) == 2 for row in handles) and len({h for row in handles for h in row}) == 14
    x, res, w = torch.randn(6, 32), torch.randn(6, 32), torch.rand(32) + 0.5
    out = torch.empty(6, 32)
    rows = [torch.tensor(r, dtype=torch.int64) for r in handles]
    ret = T.trtllm_custom_all_reduce(x, out, 2, 1, 6, T.AllReduceFusionOp.NONE, 0, 0, True, 1, rows[0], rows[2], rows[3], None, None, None, None, None, None,
                                     rows[4], rows[5], rows[6])
    assert ret is None
    torch.testing.assert_close(out, 2 * x)
    res0, bias, mid = res.clone(), torch.randn(32), torch.empty(6, 32)
    T.trtllm_custom_all_reduce(x, out, 2, 1, 6, T.AllReduceFusionOp.RESIDUAL_RMS_NORM, 0, 0, True, 2, handles[1], None, None, bias, res, w, None, 1e-5, mid)
    pre = res0 + 2 * (x + bias)
    torch.testing.assert_close(mid, pre)
    torch.testing.assert_close(out, pre * torch.rsqrt(pre.pow(2).mean(-1, keepdim=True) + 1e-5) * w)
    assert torch.equal(res, res0)                                     # the caller's residual is an input
    T.trtllm_custom_all_reduce(x, out, 2, 1, 6, T.AllReduceFusionOp.NONE, 0, 0, True, 3, handles)       # the whole table resolves too
    with pytest.raises(ValueError):
        T.trtllm_custom_all_reduce(x, out, 4, 1, 6, T.AllReduceFusionOp.NONE, 0, 0, True, 1, rows[0])
    with pytest.raises(NotImplementedError):
        T.trtllm_custom_all_reduce(x, out, 2, 1, 6, T.AllReduceFusionOp.RESIDUAL_RMS_NORM_QUANT_FP8, 0, 0, True, 1, rows[0])
    ws = handles.workspace
    T.trtllm_destroy_ipc_workspace_for_all_reduce(handles)
    assert ws.destroyed
    with pytest.raises(ValueError):
        T.trtllm_custom_all_reduce(x, out, 2, 1, 6, T.AllReduceFusionOp.NONE, 0, 0, True, 1, rows[0])
    # the fusion-era creator keeps returning (handles, workspace[, metadata]) and accepts the reference's trailing arguments
    h, w2 = T.trtllm_create_ipc_workspace_for_all_reduce_fusion(0, 2, 64, 32, False, None, False, None, False)
    assert h == [w2] and w2.is_buffer_size_sufficient(2, 64, 32, torch.bfloat16) and not w2.is_buffer_size_sufficient(2, 65, 32, torch.bfloat16)
    assert not w2.is_buffer_size_sufficient(4, 8, 32, torch.bfloat16)
