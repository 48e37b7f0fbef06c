"""Views of raw device memory as tensors (reference flashinfer/comm/dlpack_utils.py ``pack_strided_memory`` :191, which builds a DLPack
capsule by hand; the CUDA array interface expresses the same strided view in a few lines)."""
from __future__ import annotations

import torch


def pack_strided_memory(ptr: int, segment_size: int, segment_stride: int, num_segments: int, dtype: torch.dtype, dev_id):
    """Reference dlpack_utils.py:191: view ``num_segments`` segments of ``segment_size`` bytes, ``segment_stride`` bytes apart,
    starting at raw device address ``ptr`` as a ``[num_segments, segment_size / itemsize]`` tensor (no copy)."""
    esz = torch.empty(0, dtype=dtype).element_size()

    class _Raw:
        def __init__(self):
            self.__cuda_array_interface__ = {
                "shape": (num_segments, segment_size // esz), "strides": (segment_stride, esz),
                "typestr": {1: "|u1", 2: "<u2", 4: "<u4", 8: "<u8"}[esz], "data": (int(ptr), False), "version": 3}

    dev = dev_id if isinstance(dev_id, torch.device) else torch.device("cuda", int(dev_id))
    t = torch.as_tensor(_Raw(), device=dev)
    return t.view(dtype) if t.dtype != dtype else t
