"""``fi_trace(func_or_method, save_dir=None, **kwargs)``: the benchmark-definition dict of one API call, without running it.

Parity: reference flashinfer/fi_trace.py:238-285 (user API), :88-97 (deprecated registration helpers)."""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional

from .trace import TraceTemplate, registered_templates  # noqa: F401
from .trace import fi_trace as _decorator
from .trace.bindings import template_of


def fi_trace(func_or_method: Callable, save_dir: Optional[str] = None, **kwargs: Any) -> Dict[str, Any]:
    """``fi_trace(flashinfer_b200.rmsnorm, input=x, weight=w)`` -> definition dict (also written under ``save_dir`` or
    ``$FLASHINFER_TRACE_DUMP_DIR`` when given).  Works on functions, bound methods and ``Class.method``.

    Passing a :class:`TraceTemplate` instead of a function returns the binding decorator (``@fi_trace(template)``)."""
    if isinstance(func_or_method, TraceTemplate):
        return _decorator(func_or_method, save_dir)
    tpl = template_of(func_or_method)
    if tpl is None:
        name = getattr(getattr(func_or_method, "__func__", func_or_method), "__qualname__", repr(func_or_method))
        raise ValueError(f"no trace template is bound to '{name}' (see flashinfer_b200.trace.BINDINGS)")
    return tpl.dump(kwargs, save_dir) if save_dir else tpl.build_fi_trace_fn()(**kwargs)


_LEGACY_REGISTRY: Dict[str, Any] = {}


def register_fi_trace(qualname: str, spec) -> None:
    """Deprecated in the reference (fi_trace.py:88): attach a legacy trace spec to a function name."""
    _LEGACY_REGISTRY[qualname] = spec


def build_fi_trace_fn(spec):
    """Deprecated in the reference (fi_trace.py:97): the trace callable of a template / legacy spec."""
    if isinstance(spec, TraceTemplate):
        return spec.build_fi_trace_fn()
    if hasattr(spec, "build_fi_trace_fn"):
        return spec.build_fi_trace_fn()
    return lambda **kw: {"spec": repr(spec), **kw}
