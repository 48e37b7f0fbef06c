"""Module path of the reference (flashinfer/comm/torch_symmetric_memory.py): symmetric allocations (implementation: symm.py)."""
from .symm import SymmetricHeap  # noqa: F401


def alloc_symm_buffer_bytes(nbytes: int, group=None):
    """A symmetric uint8 buffer of ``nbytes`` on every rank: returns ``(local view, heap)``; ``heap.peer_ptrs`` / ``heap.mc_ptr`` address it."""
    heap = SymmetricHeap(group, nbytes)
    view, _ = heap.alloc(nbytes)
    return view, heap
