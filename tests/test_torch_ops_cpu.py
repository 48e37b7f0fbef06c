"""torch.library custom ops + fake kernels: torch.compile(fullgraph=True) traces through library calls without graph breaks and
torch.library.opcheck validates schema / fake kernel / functionalisation (reference: register_custom_op / register_fake_op,
flashinfer/utils.py:325-376; tests/conftest.py torch.compile mode)."""
import pytest
import torch

import flashinfer_b200 as fi
import flashinfer_b200.torch_ops as tops


def test_namespace_and_registry():
    for n in tops.REGISTERED:
        assert hasattr(torch.ops.flashinfer_b200, n), n


def _mlp(x, res, wn, wgu, wd):
    torch.ops.flashinfer_b200.fused_add_rmsnorm(x, res, wn, 1e-5)
    h = torch.ops.flashinfer_b200.linear(x, wgu)
    a = torch.ops.flashinfer_b200.silu_and_mul(h)
    return torch.ops.flashinfer_b200.linear(a, wd)


def test_compile_fullgraph_matches_eager():
    torch.manual_seed(0)
    x, res = torch.randn(6, 64), torch.randn(6, 64)
    wn, wgu, wd = torch.rand(64) + 0.5, torch.randn(256, 64) / 8, torch.randn(64, 128) / 11
    ref = _mlp(x.clone(), res.clone(), wn, wgu, wd)
    compiled = torch.compile(_mlp, fullgraph=True, backend="aot_eager")
    xc, rc = x.clone(), res.clone()
    got = compiled(xc, rc, wn, wgu, wd)
    torch.testing.assert_close(got, ref)
    # the mutation of (x, res) is visible to the caller of the compiled function
    xe, re_ = x.clone(), res.clone()
    _mlp(xe, re_, wn, wgu, wd)
    torch.testing.assert_close(xc, xe)
    torch.testing.assert_close(rc, re_)


@pytest.mark.parametrize("op,args", [
    ("rmsnorm", lambda: (torch.randn(4, 32), torch.rand(32), 1e-6)),
    ("silu_and_mul", lambda: (torch.randn(3, 16),)),
    ("linear", lambda: (torch.randn(5, 32), torch.randn(8, 32), None)),
    ("merge_state", lambda: (torch.randn(3, 2, 8), torch.randn(3, 2), torch.randn(3, 2, 8), torch.randn(3, 2))),
    ("top_k_renorm_probs", lambda: (torch.softmax(torch.randn(2, 50), -1), 5)),
])
def test_opcheck(op, args):
    torch.library.opcheck(getattr(torch.ops.flashinfer_b200, op).default, args(),
                          test_utils=("test_schema", "test_faketensor"))


def test_reference_named_helpers_exist():
    from flashinfer_b200 import utils

    assert callable(utils.register_custom_op) and callable(utils.register_fake_op)
