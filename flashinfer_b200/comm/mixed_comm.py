"""Reference flashinfer/comm/mixed_comm.py: NVLS collectives behind one handler (implementation in collectives.py)."""
from .collectives import MixedCommHandler, MixedCommMode, MixedCommOp, run_mixed_comm  # noqa: F401
