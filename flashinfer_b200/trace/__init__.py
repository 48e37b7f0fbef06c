"""Parity: reference flashinfer/trace (TraceTemplate, per-op templates) + flashinfer/fi_trace.py."""
from .template import Const, Scalar, TemplateDispatch, Tensor, TraceTemplate, Var, concrete_templates, dump_dir, fi_trace, registered_templates, traced  # noqa: F401
from . import templates  # noqa: F401
from .bindings import BINDINGS, FLAT_BINDINGS, attach, disable, enable, template_of  # noqa: F401
