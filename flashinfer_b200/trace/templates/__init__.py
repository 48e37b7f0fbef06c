"""Op templates, one module per category (reference flashinfer/trace/templates/*.py)."""
from . import activation, norm, rope, sampling  # noqa: F401
from ._legacy import *  # noqa: F401,F403
from .activation import *  # noqa: F401,F403
from .norm import *  # noqa: F401,F403
from .rope import *  # noqa: F401,F403
from .sampling import *  # noqa: F401,F403
