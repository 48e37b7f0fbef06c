"""Module path of the reference (flashinfer/gdn_prefill.py); implementation: gdn.py."""
from .gdn import chunk_gated_delta_rule  # noqa: F401
