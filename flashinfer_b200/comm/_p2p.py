"""Uneven all-to-all built from isend / irecv (works on every backend, including gloo)."""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def all_to_all_uneven(recvs: List[torch.Tensor], sends: List[torch.Tensor], group=None) -> None:
    group = group if group is not None else dist.group.WORLD
    me = dist.get_rank(group)
    ranks = dist.get_process_group_ranks(group)
    recvs[me].copy_(sends[me])
    ops = []
    for i, r in enumerate(ranks):
        if i == me:
            continue
        if sends[i].numel():
            ops.append(dist.P2POp(dist.isend, sends[i], r, group))
        if recvs[i].numel():
            ops.append(dist.P2POp(dist.irecv, recvs[i], r, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
