"""Parity: reference flashinfer/fi_trace.py — re-export of the tracing decorator and template registry."""
from .trace import TraceTemplate, fi_trace, registered_templates  # noqa: F401
