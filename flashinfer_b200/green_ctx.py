"""SM partitioning with CUDA green contexts.  Parity: reference flashinfer/green_ctx.py:34-295
(split_device_green_ctx / split_device_green_ctx_by_sm_count returning streams bound to disjoint SM sets).

On B200 the SM-resource split granularity is 8 SMs (minimum 8); 148 SMs = 18 groups of 8 + 4 remainder.
Used by POD-style co-scheduling (prefill on one partition, decode on the other).
"""
from __future__ import annotations

from typing import List, Tuple

import torch


def _drv():
    try:
        from cuda.bindings import driver  # cuda-python >= 12.8
    except ImportError:  # pragma: no cover
        from cuda import cuda as driver
    return driver


def _check(res):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError(f"CUDA driver error {err}")
    return res[1] if len(res) == 2 else res[1:]


def get_sm_count_constraint(major: int = 10, minor: int = 0) -> Tuple[int, int]:
    """(minimum partition size, alignment) of the SM resource split."""
    if major >= 9:
        return 8, 8
    if major == 8:
        return 4, 2
    return 2, 2


def round_up_sm_count(n: int, device=None) -> int:
    mn, al = get_sm_count_constraint(*torch.cuda.get_device_capability(device))
    return max(mn, (n + al - 1) // al * al)


def _sm_resource(dev_index: int):
    drv = _drv()
    dev = _check(drv.cuDeviceGet(dev_index))
    return drv, dev, _check(drv.cuDeviceGetDevResource(dev, drv.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))


def _streams_from_resources(drv, dev, resources) -> List[torch.cuda.Stream]:
    streams = []
    for r in resources:
        desc = _check(drv.cuDevResourceGenerateDesc([r], 1))
        gctx = _check(drv.cuGreenCtxCreate(desc, dev, drv.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
        st = _check(drv.cuGreenCtxStreamCreate(gctx, drv.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
        streams.append(torch.cuda.get_stream_from_external(int(st), torch.device("cuda", torch.cuda.current_device())))
    return streams


def split_device_green_ctx(dev: torch.device, num_groups: int, min_count: int):
    """``num_groups`` partitions of >= ``min_count`` SMs each plus one partition with the remaining SMs.
    Returns ``(streams, resources)``; work submitted to ``streams[i]`` only runs on partition ``i``."""
    torch.cuda.init()
    idx = torch.device(dev).index or 0
    drv, cdev, sm = _sm_resource(idx)
    min_count = round_up_sm_count(min_count, idx)
    res = drv.cuDevSmResourceSplitByCount(num_groups, sm, 0, min_count)
    err, groups, n, remaining = res
    if int(err) != 0:
        raise RuntimeError(f"cuDevSmResourceSplitByCount failed: {err}")
    resources = list(groups[:n]) + [remaining]
    return _streams_from_resources(drv, cdev, resources), resources


def split_device_green_ctx_by_sm_count(dev: torch.device, sm_counts: List[int]):
    """One partition per requested SM count (each rounded up to the split granularity) + the remainder."""
    torch.cuda.init()
    idx = torch.device(dev).index or 0
    drv, cdev, sm = _sm_resource(idx)
    resources = []
    cur = sm
    for c in sm_counts:
        c = round_up_sm_count(c, idx)
        err, groups, n, remaining = drv.cuDevSmResourceSplitByCount(1, cur, 0, c)
        if int(err) != 0 or n < 1:
            raise RuntimeError(f"cannot carve {c} SMs out of the remaining resource: {err}")
        resources.append(groups[0])
        cur = remaining
    resources.append(cur)
    return _streams_from_resources(drv, cdev, resources), resources


def get_sm_count(resource) -> int:
    return int(resource.sm.smCount)


# ------------------------------------------------------------------ building blocks under the reference's names (green_ctx.py:47-125)
def get_cudevice(dev: torch.device):
    torch.cuda.init()
    return _check(_drv().cuDeviceGet(torch.device(dev).index or 0))


def get_device_resource(cu_dev):
    drv = _drv()
    return _check(drv.cuDeviceGetDevResource(cu_dev, drv.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))


def split_resource(resource, num_groups: int, min_count: int):
    """``(groups, remaining)`` of ``cuDevSmResourceSplitByCount``."""
    err, groups, n, remaining = _drv().cuDevSmResourceSplitByCount(num_groups, resource, 0, min_count)
    if int(err) != 0:
        raise RuntimeError(f"cuDevSmResourceSplitByCount failed: {err}")
    return list(groups[:n]), remaining


def split_resource_by_sm_count(cu_dev, resource, sm_counts: List[int]):
    results, cur = [], resource
    for c in sm_counts:
        got, cur = split_resource(cur, 1, c)
        results.extend(got)
    return results, cur


def create_green_ctx_streams(cu_dev, resources) -> List[torch.cuda.Stream]:
    return _streams_from_resources(_drv(), cu_dev, resources)
