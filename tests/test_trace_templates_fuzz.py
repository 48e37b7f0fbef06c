"""Every bound fi_trace template against its API over several seeds and batch-like sizes (1, 2, 9) - the shapes where host-side
planning / masking / padding logic tends to break (single request, odd counts).  CPU paths; ~1100 (template, seed, size) cases."""
import inspect

import torch

import test_trace_templates as T
from flashinfer_b200.trace import Var


def _sizes(tpl, var_size):
    accepted = inspect.signature(tpl.init).parameters
    sizes = dict(tpl.test_sizes or {})
    for a in tpl.axes:
        if a.name in accepted:
            if isinstance(a, Var) and a.name not in (tpl.test_sizes or {}):
                sizes[a.name] = var_size
            elif a.name not in sizes:
                sizes[a.name] = 128 if ("size" in a.name or "dim" in a.name) else 4
    return sizes


def _check(tpl, api, kwargs):
    ref_in = {k: (v.clone() if isinstance(v, torch.Tensor) else tuple(t.clone() for t in v) if isinstance(v, tuple) else v) for k, v in kwargs.items()}
    expect = tpl.run_reference(ref_in)
    expect = list(expect) if isinstance(expect, (tuple, list)) else [expect]
    got = tpl.collect_outputs(api(**kwargs), kwargs)
    if tpl.compare is not None:
        tpl.compare(got, expect, ref_in)
        return
    tol = T.TOLERANCE[tpl.tolerance]
    for g, e in zip(got, expect):
        if tol == "cos":
            assert torch.nn.functional.cosine_similarity(g.float().flatten(), e.float().flatten(), dim=0) > 0.99
        elif tol is None:
            assert e[torch.arange(g.numel()), g.long()].all()
        elif tol["atol"] == 0 and tol["rtol"] == 0:
            assert torch.equal(g, e)
        else:
            torch.testing.assert_close(g.float(), e.float(), **tol)


def test_templates_over_seeds_and_sizes():
    failures, cases = [], 0
    for (m, p, tpl), tid in zip(T.BINDINGS, T._IDS):
        if tpl.init is None:
            continue
        api = T._api(m, p)
        for seed in (2, 3, 5):
            for var_size in (1, 2, 9):
                try:
                    _check(tpl, api, tpl.make_inputs(device="cpu", seed=seed, **_sizes(tpl, var_size)))
                    cases += 1
                except Exception as exc:  # noqa: BLE001
                    failures.append(f"{tid} seed={seed} size={var_size}: {type(exc).__name__}: {str(exc)[:160]}")
    assert not failures, "\n".join(failures[:20])
    assert cases > 1000
