"""Module path of the reference (flashinfer/gdn_prefill.py); implementation: gdn.py."""
from .gdn import chunk_gated_delta_rule  # noqa: F401


from . import jit as _jit_acc  # noqa: E402

get_gdn_prefill_module = _jit_acc.module_accessor("ssm")
