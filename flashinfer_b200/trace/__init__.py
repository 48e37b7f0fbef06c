"""Parity: reference flashinfer/trace (TraceTemplate, per-op templates) + flashinfer/fi_trace.py."""
from .template import Const, Scalar, Tensor, TraceTemplate, Var, fi_trace, registered_templates  # noqa: F401
from . import templates  # noqa: F401
