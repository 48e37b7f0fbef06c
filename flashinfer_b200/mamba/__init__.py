"""Mamba state-space ops.  Parity: reference flashinfer/mamba (selective_state_update.py:104)."""
from .selective_state_update import selective_state_update, selective_state_update_ref  # noqa: F401
