"""Module path of the reference (flashinfer/comm/workspace_base.py)."""
from .compat import AllReduceFusionWorkspace  # noqa: F401
