"""Expert-parallel dispatch / combine backends.

Parity: reference flashinfer/moe_ep/__init__.py:1-128 (import-time probe of the third-party NCCL-EP / NIXL-EP plugin
libraries, ``MoEEpNotBuiltError``, ``available_backends``).  Neither plugin exists in this image and this framework does
not wrap them: its expert-parallel transport is the in-tree NVLink symmetric-heap kernel pair (``csrc/comm/moe_a2a.cu``:
push dispatch with in-kernel completion handshake, pull combine with fused reduction) behind
:class:`flashinfer_b200.comm.MoeAlltoAll`.  This module keeps the reference's probe surface and adds that native backend
under the name ``"nvlink_a2a"`` so code that asks "which EP backends can I use" gets a truthful answer.
"""
from __future__ import annotations

from typing import List

__all__ = ["MoEEpNotBuiltError", "have_nccl_ep", "have_nixl_ep", "have_nvlink_a2a", "available_backends", "create_fleet"]


class MoEEpNotBuiltError(RuntimeError):
    """An EP backend was requested whose native library is not part of this build."""


def have_nccl_ep() -> bool:
    """NCCL-EP plugin (``libnccl_ep.so``) present?  Never in this build (no third-party EP plugins are vendored)."""
    return False


def have_nixl_ep() -> bool:
    """NIXL-EP plugin (``nixl_ep_cpp*.so``) present?  Never in this build."""
    return False


def have_nvlink_a2a() -> bool:
    """The in-tree dispatch / combine kernels (always buildable: one .cu, no dependencies beyond CUDA)."""
    from .. import jit

    return "comm_alltoall" in jit.REGISTRY


def available_backends() -> List[str]:
    out = []
    if have_nccl_ep():
        out.append("nccl_ep")
    if have_nixl_ep():
        out.append("nixl_ep")
    if have_nvlink_a2a():
        out.append("nvlink_a2a")
    return out


def _require_built(backend: str) -> None:
    probes = {"nccl_ep": have_nccl_ep, "nixl_ep": have_nixl_ep, "nvlink_a2a": have_nvlink_a2a}
    if backend not in probes:
        raise ValueError(f"unknown moe_ep backend {backend!r}; expected one of {', '.join(probes)}")
    if not probes[backend]():
        raise MoEEpNotBuiltError(f"moe_ep backend {backend!r} is not part of this build; use 'nvlink_a2a' "
                                 "(flashinfer_b200.comm.MoeAlltoAll) on a single NVLink domain")


def create_fleet(mapping, max_num_tokens: int, top_k: int, num_experts: int, hidden_size: int,
                 backend: str = "nvlink_a2a", **kwargs):
    """Build the dispatch / combine object for one EP group.  ``backend='nvlink_a2a'`` returns a
    :class:`~flashinfer_b200.comm.MoeAlltoAll`; the plugin backends raise :class:`MoEEpNotBuiltError`."""
    _require_built(backend)
    from ..comm.trtllm_moe_alltoall import MoeAlltoAll

    return MoeAlltoAll(mapping, max_num_tokens, top_k, num_experts, hidden_size=hidden_size, **kwargs)
