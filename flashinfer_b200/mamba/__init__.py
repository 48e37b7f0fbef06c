"""Mamba state-space ops.  Parity: reference flashinfer/mamba (selective_state_update.py:104, ssd_combined.py:250)."""
from .selective_state_update import selective_state_update, selective_state_update_ref  # noqa: F401
from .ssd_combined import SSDCombined, chunk_cumsum_fwd, ssd_combined_fwd, ssd_reference  # noqa: F401
