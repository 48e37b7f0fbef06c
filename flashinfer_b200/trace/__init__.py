"""Parity: reference flashinfer/trace (TraceTemplate, per-op templates) + flashinfer/fi_trace.py."""
from .template import Const, Scalar, Tensor, TraceTemplate, Var, dump_dir, fi_trace, registered_templates, traced  # noqa: F401
from . import templates  # noqa: F401
from .bindings import BINDINGS, attach, disable, enable, template_of  # noqa: F401
