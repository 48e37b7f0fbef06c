"""Op templates, one module per category (reference flashinfer/trace/templates/*.py)."""
from . import activation, attention, attention_more, cascade, gdn, gemm, misc, moe, moe_trtllm, more_ops, norm, page, quantize, rope, round2, sampling  # noqa: F401
from .comm import *  # noqa: F401,F403
from .activation import *  # noqa: F401,F403
from .norm import *  # noqa: F401,F403
from .rope import *  # noqa: F401,F403
from .sampling import *  # noqa: F401,F403
from .cascade import *  # noqa: F401,F403
from .page import *  # noqa: F401,F403
from .attention import *  # noqa: F401,F403
from .gemm import *  # noqa: F401,F403
from .quantize import *  # noqa: F401,F403
from .moe import *  # noqa: F401,F403
from .moe_trtllm import *  # noqa: F401,F403
from .misc import *  # noqa: F401,F403
from .gdn import *  # noqa: F401,F403
from .round2 import *  # noqa: F401,F403
from .attention_more import *  # noqa: F401,F403
from .more_ops import *  # noqa: F401,F403
