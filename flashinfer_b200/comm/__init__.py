"""Communication kernels over NVLink 5 / NVSwitch (symmetric heap + in-kernel collectives)."""
from .allreduce import TPCommunicator  # noqa: F401
from .symm import SymmetricHeap  # noqa: F401
from .mapping import Mapping  # noqa: F401,E402
