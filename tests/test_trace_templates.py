"""fi_trace templates: every bound template must (1) name real parameters of the API it is bound to, (2) produce a
well-formed definition, and (3) carry a reference implementation that agrees with the API on inputs built by its own
``init`` (reference tests/trace/test_*_reference_correctness.py + test_fi_trace_template_consistency.py, as one generic
test).  Runs on CPU through the ops' eager paths; the same harness is device-agnostic."""
import inspect
import json
import os

import pytest
import torch

import flashinfer_b200 as fi
from flashinfer_b200.trace import FLAT_BINDINGS as BINDINGS, Const, Scalar, Tensor, Var
from flashinfer_b200.trace.bindings import _resolve

# FIB200_TRACE_TEST_DEVICE=cuda runs the same checks against the native kernels (opt-in: not part of the gpu-marked suite yet)
DEVICE = os.environ.get("FIB200_TRACE_TEST_DEVICE", "cpu")

# tolerance classes (reference tests/trace/reference_correctness_standards.md)
TOLERANCE = {
    "exact": dict(atol=0.0, rtol=0.0),
    "bf16_norm": dict(atol=2e-2, rtol=2e-2),
    "bf16": dict(atol=3e-2, rtol=3e-2),
    "fp32": dict(atol=1e-5, rtol=1e-5),
    "fp8_quant": dict(atol=0.0, rtol=0.13),     # one e4m3 step
    "cos": "cos",                               # GEMMs: cosine similarity > 0.99 plus a loose elementwise bound
    "support": None,                            # stochastic: the reference is the mask of tokens that may be emitted
}

_IDS = [f"{m}.{p}" + (f"[{t.name_fmt.split('{')[0].rstrip('_e')}]" if sum(1 for b in BINDINGS if b[:2] == (m, p)) > 1 else "") for m, p, t in BINDINGS]


def _api(mod, path):
    owner, attr = _resolve(mod, path)
    return getattr(owner, attr)


@pytest.mark.parametrize("mod,path,tpl", BINDINGS, ids=_IDS)
def test_template_names_real_parameters(mod, path, tpl):
    params = inspect.signature(_api(mod, path)).parameters
    for spec in tpl.inputs:
        if spec.source.startswith("self."):
            assert "self" in params                   # plan()-time state: checked on a live instance below
            continue
        assert spec.source in params, f"{tpl.key}: input '{spec.name}' reads API parameter '{spec.source}' which does not exist"
    for spec in tpl.outputs:
        if spec.param is not None:
            assert spec.param in params
    axis_names = {a.name for a in tpl.axes}
    assert all(isinstance(a, (Const, Var)) for a in tpl.axes)
    derived = set()
    for spec in list(tpl.inputs) + list(tpl.outputs):
        if isinstance(spec, Tensor):
            derived |= set(spec.axes) - axis_names
    # axes used in shapes but not declared must be explained by a constraint
    for ax in derived:
        assert any(ax in c for c in tpl.constraints), f"{tpl.key}: axis '{ax}' is neither declared nor constrained"
    if tpl.reference is not None:
        ref_params = set(inspect.signature(tpl.reference).parameters)
        assert ref_params <= {s.name for s in tpl.inputs}, f"{tpl.key}: reference takes arguments the template does not list"


@pytest.mark.parametrize("mod,path,tpl", [b for b in BINDINGS if b[2].init is not None], ids=[i for i, b in zip(_IDS, BINDINGS) if b[2].init is not None])
def test_reference_matches_api(mod, path, tpl):
    api = _api(mod, path)
    sizes = dict(tpl.test_sizes or {})
    accepted = inspect.signature(tpl.init).parameters
    for a in tpl.axes:
        if a.name in accepted and a.name not in sizes:
            sizes[a.name] = 5 if isinstance(a, Var) else (128 if "size" in a.name or "dim" in a.name else 4)
    kwargs = tpl.make_inputs(device=DEVICE, seed=1, **sizes)
    ref_in = {k: (v.clone() if isinstance(v, torch.Tensor) else tuple(t.clone() for t in v) if isinstance(v, tuple) else v)
              for k, v in kwargs.items()}
    from flashinfer_b200.trace.template import _pick

    for spec in tpl.inputs:
        assert spec.optional or _pick(kwargs, spec) is not None, f"{tpl.key}: input '{spec.name}' ({spec.source}) did not resolve"
    expect = tpl.run_reference(ref_in)
    expect = list(expect) if isinstance(expect, (tuple, list)) else [expect]
    got = tpl.collect_outputs(api(**kwargs), kwargs)
    assert len(got) >= len(expect) > 0
    tol = TOLERANCE[tpl.tolerance]
    if tpl.compare is not None:
        tpl.compare(got, expect, ref_in)
        got, expect = [], []
    for spec, g, e in zip(tpl.outputs, got, expect):
        assert isinstance(g, torch.Tensor), f"{tpl.key}: output '{spec.name}' missing"
        if tol == "cos":
            assert g.shape == e.shape
            cos = torch.nn.functional.cosine_similarity(g.float().flatten(), e.float().flatten(), dim=0)
            assert cos > 0.99, f"{tpl.key}/{spec.name}: cosine similarity {float(cos):.4f}"
            torch.testing.assert_close(g.float(), e.float(), atol=0.1 * float(e.float().abs().max()), rtol=0.05)
            continue
        if tol is None:
            assert e.dtype == torch.bool and g.shape == e.shape[:1]
            assert e[torch.arange(g.numel(), device=g.device), g.long()].all(), f"{tpl.key}: sampled a token outside the filtered support"
            if len(tpl.inputs) > 1:                          # a filtered sampler: the filter must actually remove something
                assert not e.all(), f"{tpl.key}: degenerate test, the filter keeps everything"
            continue
        assert g.shape == e.shape, f"{tpl.key}: output '{spec.name}' shape {tuple(g.shape)} vs reference {tuple(e.shape)}"
        if tol["atol"] == 0.0 and tol["rtol"] == 0.0:
            assert torch.equal(g, e), f"{tpl.key}: output '{spec.name}' differs"
        else:
            torch.testing.assert_close(g.float(), e.float(), **tol, msg=lambda m: f"{tpl.key}/{spec.name}: {m}")
    # the definition of this very call is well formed and names the const axes
    d = api.fi_trace(**kwargs) if hasattr(api, "fi_trace") else tpl.definition(kwargs)
    json.dumps(d)
    assert d["op_type"] == tpl.op_type and "{" not in d["name"]
    for a in tpl.axes:
        if isinstance(a, Const):
            assert d["axes"][a.name]["value"] is not None, f"{tpl.key}: const axis '{a.name}' not resolved from the call"
    assert all(v["dtype"] != "unknown" for k, v in d["inputs"].items() if not v.get("optional"))


def test_enable_disable_swaps_wrappers_and_dumps_once(tmp_path):
    x, w = torch.randn(4, 256).bfloat16(), torch.ones(256).bfloat16()
    plain = fi.rmsnorm
    assert fi.trace.dump_dir() is None and not hasattr(plain, "__wrapped_untraced__")
    fi.trace.enable(str(tmp_path))
    try:
        assert fi.rmsnorm is not plain and fi.norm.rmsnorm is fi.rmsnorm and fi.rmsnorm.__wrapped_untraced__ is plain
        fi.rmsnorm(x, w)
        f = tmp_path / "rmsnorm" / "rmsnorm_h256.json"
        stamp = f.stat().st_mtime_ns
        fi.rmsnorm(x, w)                                   # same definition: not rewritten
        assert f.stat().st_mtime_ns == stamp
        fi.rmsnorm(torch.randn(2, 512).bfloat16(), torch.ones(512).bfloat16())
        assert (tmp_path / "rmsnorm" / "rmsnorm_h512.json").exists()
        d = json.loads(f.read_text())
        assert d["axes"]["hidden_size"] == {"type": "const", "value": 256} and "def _rmsnorm_reference" in d["reference"]
        ns = {"torch": torch}
        exec(d["reference"], ns)                           # the embedded source runs on its own
        torch.testing.assert_close(ns["_rmsnorm_reference"](x, w).float(), plain(x, w).float(), atol=2e-2, rtol=2e-2)
    finally:
        fi.trace.disable()
    assert fi.rmsnorm is plain and fi.norm.rmsnorm is plain


def test_fi_trace_user_api(tmp_path):
    x = torch.randn(3, 512).bfloat16()
    d = fi.fi_trace(fi.silu_and_mul, input=x)
    assert d["name"] == "silu_and_mul_h256" and d["tags"][0] == "fi_api:flashinfer_b200.activation.silu_and_mul"
    fi.fi_trace(fi.silu_and_mul, save_dir=str(tmp_path), input=x)
    assert (tmp_path / "activation" / "silu_and_mul_h256.json").exists()
    with pytest.raises(ValueError):
        fi.fi_trace(torch.relu, input=x)


def test_mxfp4_template_against_fp4_quantize():
    """The MXFP4 flavour shares fp4_quantize with NVFP4 (only one template can be bound to a function)."""
    from flashinfer_b200.trace.templates import mxfp4_quantize_trace as tpl

    kw = tpl.make_inputs(device="cpu", seed=3, M=5, K=128)
    tpl.compare(list(fi.fp4_quantize(**kw)), list(tpl.run_reference(kw)), kw)
    assert tpl.definition(kw)["name"] == "mxfp4_quantize_k128"


def test_enable_wraps_wrapper_methods(tmp_path):
    from flashinfer_b200.trace.templates import gqa_paged_decode_trace as tpl

    kw = tpl.make_inputs(device="cpu", seed=0, batch_size=2, num_qo_heads=4, num_kv_heads=2, head_dim=64, page_size=4)
    w = kw.pop("self")
    plain = fi.BatchDecodeWithPagedKVCacheWrapper.run
    fi.trace.enable(str(tmp_path))
    try:
        assert fi.BatchDecodeWithPagedKVCacheWrapper.run is not plain
        w.run(kw["q"], kw["paged_kv_cache"])
        files = list((tmp_path / "gqa_paged").iterdir())
        assert [f.name for f in files] == ["gqa_paged_decode_h4_kv2_d64_ps4.json"]
        d = json.loads(files[0].read_text())
        assert d["inputs"]["kv_indptr"]["dtype"] == "int32" and d["axes"]["page_size"]["value"] == 4
        assert d["tags"][0] == "fi_api:flashinfer_b200.decode.BatchDecodeWithPagedKVCacheWrapper.run"
    finally:
        fi.trace.disable()
    assert fi.BatchDecodeWithPagedKVCacheWrapper.run is plain
    assert fi.fi_trace(w.run, q=kw["q"], paged_kv_cache=kw["paged_kv_cache"])["axes"]["head_dim"]["value"] == 64


def test_every_category_of_the_reference_has_templates():
    from flashinfer_b200.trace import registered_templates

    cats = {k.split(":")[0] for k in registered_templates()}
    assert {"rmsnorm", "layernorm", "activation", "rope", "sampling", "cascade", "page", "gemm", "quantize", "gqa_single", "gqa_paged",
            "gqa_ragged", "mla_paged", "moe", "comm"} <= cats
    assert len(registered_templates()) >= 60
