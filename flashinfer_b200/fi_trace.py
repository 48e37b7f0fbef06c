"""Parity: reference flashinfer/fi_trace.py — re-export of the tracing decorator and template registry."""
from .trace import TraceTemplate, fi_trace, registered_templates  # noqa: F401


_LEGACY_REGISTRY = {}


def register_fi_trace(qualname: str, spec) -> None:
    """Deprecated in the reference (fi_trace.py:88): attach a legacy trace spec to a function name."""
    _LEGACY_REGISTRY[qualname] = spec


def build_fi_trace_fn(spec):
    """Deprecated in the reference (fi_trace.py:97): build the trace callable of a template / legacy spec."""
    if hasattr(spec, "build_fi_trace_fn"):
        return spec.build_fi_trace_fn()
    if isinstance(spec, TraceTemplate):
        return lambda **kw: spec.render(**kw) if hasattr(spec, "render") else {"op_type": getattr(spec, "op_type", None), **kw}
    return lambda **kw: {"spec": repr(spec), **kw}
