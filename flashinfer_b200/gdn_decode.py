"""Module path of the reference (flashinfer/gdn_decode.py); implementation: gdn.py."""
from .gdn import gated_delta_rule_decode, gated_delta_rule_decode_pretranspose, gated_delta_rule_mtp  # noqa: F401
