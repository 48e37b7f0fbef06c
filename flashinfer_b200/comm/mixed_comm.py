"""TP x DP collectives behind one handler (reference flashinfer/comm/mixed_comm.py: ``ParallelInfo`` :143, ``MixedCommHandler`` :423,
``run_mixed_comm(op, handler, x_in, x_out, mode)`` :1422).

Topology model of the reference: ``world = inter_size x local_size`` ranks (nodes x GPUs per node); inside a node the ranks split into
``local_dp_size`` data-parallel replicas of ``local_tp_size`` tensor-parallel ranks (TP varies fastest), and the same split exists
across nodes.  The five ops:

=========================  ==========================================================================================
``ALLREDUCE``              sum over the TP group
``ALLGATHER``              concatenate over the DP group (ordered by dp rank) -> ``[dp_size * n, ...]``
``REDUCESCATTER``          ``[dp_size * n, ...]`` summed over the DP group, rank keeps the chunk of its dp rank
``ALLREDUCE_ALLGATHER``    TP all-reduce, then DP all-gather
``REDUCESCATTER_ALLREDUCE``  DP reduce-scatter, then TP all-reduce
=========================  ==========================================================================================

The reference fuses each op into one kernel over virtual-memory peers inside a node and NVSHMEM across nodes.  Here one NVSwitch domain
is the fabric: an op whose group is the whole world runs on the NVLS kernels (:class:`~flashinfer_b200.comm.allreduce.TPCommunicator`
for the all-reduce, :class:`~flashinfer_b200.comm.collectives.NVLSCollectives` for gather / scatter); a genuine TP x DP split runs the
two legs on ``torch.distributed`` sub-groups (NCCL on GPUs).  The reference's mode enum is kept: the ``FUSED_*`` members select the
NVLS path when it applies, ``NCCL_*`` force ``torch.distributed``, ``AUTOTUNE`` = fused when it applies."""
from __future__ import annotations

import enum
from typing import List, Optional

import torch
import torch.distributed as dist

from .. import jit as _jit_acc


class MixedCommOp(enum.IntEnum):
    ALLREDUCE = 0
    ALLGATHER = enum.auto()
    REDUCESCATTER = enum.auto()
    ALLREDUCE_ALLGATHER = enum.auto()
    REDUCESCATTER_ALLREDUCE = enum.auto()


class MixedCommMode(enum.IntEnum):
    FUSED_OPT_WAITS_MC = 0
    FUSED_OPT_WAITS_UC = enum.auto()
    FUSED_OPT_BYTES1_MC = enum.auto()
    FUSED_OPT_BYTES1_UC = enum.auto()
    FUSED_OPT_BYTES2_MC = enum.auto()
    FUSED_OPT_BYTES2_UC = enum.auto()
    NCCL_ONE = enum.auto()          # one collective on the world group (pre / post-processing on the device)
    NCCL_TP_DP = enum.auto()        # one collective per axis on the TP / DP sub-groups
    AUTOTUNE = enum.auto()


_FUSED = tuple(m for m in MixedCommMode if m.name.startswith("FUSED"))
_MC = (MixedCommMode.FUSED_OPT_WAITS_MC, MixedCommMode.FUSED_OPT_BYTES1_MC, MixedCommMode.FUSED_OPT_BYTES2_MC)


def _split(total: int, tp: Optional[int], dp: Optional[int]):
    if tp is None and dp is None:
        return total, 1
    if tp is None:
        assert total % dp == 0
        return total // dp, dp
    if dp is None:
        assert total % tp == 0
        return tp, total // tp
    assert total == tp * dp
    return tp, dp


class ParallelInfo:
    """Rank arithmetic of the nodes x GPUs, TP x DP layout (TP fastest on both levels)."""

    def __init__(self, world_rank: int, world_size: int, local_rank: int, local_size: int, inter_rank: int, inter_size: int,
                 local_tp_size: Optional[int], local_dp_size: Optional[int], inter_tp_size: Optional[int], inter_dp_size: Optional[int]):
        assert world_rank == inter_rank * local_size + local_rank
        assert world_size == inter_size * local_size
        self.world_rank, self.world_size = world_rank, world_size
        self.local_rank, self.local_size = local_rank, local_size
        self.inter_rank, self.inter_size = inter_rank, inter_size
        self.local_tp_size, self.local_dp_size = _split(local_size, local_tp_size, local_dp_size)
        self.inter_tp_size, self.inter_dp_size = _split(inter_size, inter_tp_size, inter_dp_size)
        self.local_tp_rank, self.local_dp_rank = local_rank % self.local_tp_size, local_rank // self.local_tp_size
        self.inter_tp_rank, self.inter_dp_rank = inter_rank % self.inter_tp_size, inter_rank // self.inter_tp_size

    tp_rank = property(lambda s: s.local_tp_rank + s.inter_tp_rank * s.local_tp_size)
    tp_size = property(lambda s: s.local_tp_size * s.inter_tp_size)
    dp_rank = property(lambda s: s.local_dp_rank + s.inter_dp_rank * s.local_dp_size)
    dp_size = property(lambda s: s.local_dp_size * s.inter_dp_size)
    use_local_tp = property(lambda s: s.local_tp_size > 1)
    use_inter_tp = property(lambda s: s.inter_tp_size > 1)
    use_tp = property(lambda s: s.tp_size > 1)
    use_dp = property(lambda s: s.dp_size > 1)
    use_inter = property(lambda s: s.inter_size > 1)
    use_mixed = property(lambda s: s.tp_size > 1 and s.dp_size > 1)

    def get_local_full_group_local_ranks(self) -> List[int]:
        return list(range(self.local_size))

    def get_local_tp_group_local_ranks(self, local_dp_rank: Optional[int] = None) -> List[int]:
        d = self.local_dp_rank if local_dp_rank is None else local_dp_rank
        return [d * self.local_tp_size + t for t in range(self.local_tp_size)]

    def get_local_dp_group_local_ranks(self, local_tp_rank: Optional[int] = None) -> List[int]:
        t = self.local_tp_rank if local_tp_rank is None else local_tp_rank
        return [d * self.local_tp_size + t for d in range(self.local_dp_size)]

    def tp_groups(self) -> List[List[int]]:
        """World ranks of every TP group, indexed by dp rank (members ordered by tp rank)."""
        out = []
        for idp in range(self.inter_dp_size):
            for ldp in range(self.local_dp_size):
                ranks = []
                for itp in range(self.inter_tp_size):
                    node = idp * self.inter_tp_size + itp
                    ranks += [r + node * self.local_size for r in self.get_local_tp_group_local_ranks(ldp)]
                out.append(ranks)
        return out

    def dp_groups(self) -> List[List[int]]:
        """World ranks of every DP group, indexed by tp rank (members ordered by dp rank)."""
        out = []
        for itp in range(self.inter_tp_size):
            for ltp in range(self.local_tp_size):
                ranks = []
                for idp in range(self.inter_dp_size):
                    node = idp * self.inter_tp_size + itp
                    ranks += [r + node * self.local_size for r in self.get_local_dp_group_local_ranks(ltp)]
                out.append(ranks)
        return out

    def _pick(self, groups: List[List[int]], mine: int):
        made = [dist.new_group(g) for g in groups]           # collective: every rank creates every group, in the same order
        return made[mine]

    def get_local_comm_group(self):
        return self._pick([list(range(n * self.local_size, (n + 1) * self.local_size)) for n in range(self.inter_size)], self.inter_rank)

    def get_tp_comm_group(self):
        return self._pick(self.tp_groups(), self.dp_rank)

    def get_dp_comm_group(self):
        return self._pick(self.dp_groups(), self.tp_rank)


class MixedCommHandler:
    """Reference constructor (mixed_comm.py :453).  ``grid_size`` / block sizes / ``min_num_steps`` shape the reference's fused kernel,
    ``ib_enable_ibgda`` / ``should_init_nvshmem`` configure NVSHMEM: accepted for call-site compatibility, they have no counterpart on one
    NVSwitch domain.  ``max_tokens`` / ``hidden`` (extensions) size the symmetric heaps of the NVLS path."""

    def __init__(self, world_rank: int, world_size: int, local_rank: int, local_size: int, inter_rank: int, inter_size: int,
                 local_tp_size: Optional[int], local_dp_size: Optional[int], inter_tp_size: Optional[int], inter_dp_size: Optional[int],
                 dtype: torch.dtype, device: torch.device, grid_size: Optional[int] = None, max_block_size: Optional[int] = None,
                 min_block_size: int = 256, min_num_steps: int = 4, ib_enable_ibgda: bool = True, should_init_nvshmem: bool = True,
                 use_autotune: bool = True, *, max_tokens: int = 8192, hidden: int = 8192) -> None:
        assert dist.is_initialized()
        assert local_size > 1
        assert dtype in (torch.float16, torch.bfloat16, torch.float32)
        self.is_running = True
        self.para_info = ParallelInfo(world_rank, world_size, local_rank, local_size, inter_rank, inter_size, local_tp_size, local_dp_size,
                                      inter_tp_size, inter_dp_size)
        self.dtype, self.device = dtype, torch.device(device)
        self.use_autotune = use_autotune
        p = self.para_info
        self.tp_comm_group = p.get_tp_comm_group() if p.use_mixed else (dist.group.WORLD if p.use_tp else None)
        self.dp_comm_group = p.get_dp_comm_group() if p.use_mixed else (dist.group.WORLD if p.use_dp else None)
        self.valid_op_list = self.get_valid_op_list()
        self.valid_mode_list = self.get_valid_mode_list()
        self._ar = self._coll = None
        if self.device.type == "cuda" and not p.use_mixed:     # the op's group is the whole world: NVLS kernels
            from .allreduce import TPCommunicator
            from .collectives import NVLSCollectives

            esz = torch.empty((), dtype=dtype).element_size()
            if p.use_tp:
                self._ar = TPCommunicator(dist.group.WORLD, max_tokens, hidden, dtype)
            else:
                self._coll = NVLSCollectives(dist.group.WORLD, max_tokens * hidden * esz)

    # ---- reference surface
    def get_valid_op_list(self) -> List[MixedCommOp]:
        p = self.para_info
        ops = []
        if p.use_tp:
            ops.append(MixedCommOp.ALLREDUCE)
        if p.use_dp:
            ops += [MixedCommOp.ALLGATHER, MixedCommOp.REDUCESCATTER]
        if p.use_mixed:
            ops += [MixedCommOp.ALLREDUCE_ALLGATHER, MixedCommOp.REDUCESCATTER_ALLREDUCE]
        return ops

    def get_valid_mode_list(self) -> List[MixedCommMode]:
        modes = [MixedCommMode.NCCL_ONE]
        if self.para_info.use_mixed:
            modes.append(MixedCommMode.NCCL_TP_DP)
        if self.device.type == "cuda" and not self.para_info.use_mixed:
            modes = list(_FUSED) + modes
        return modes

    def select_autotune_mode(self, op: MixedCommOp, x_in: torch.Tensor) -> MixedCommMode:
        return self.valid_mode_list[0]

    def run_autotune(self) -> None:
        """The reference times every mode per op and message size; one fused path exists here, nothing to choose."""

    def shutdown(self) -> None:
        self.is_running = False
        self._ar = self._coll = None

    # ---- ops
    def run(self, op: MixedCommOp, x_in: torch.Tensor, x_out: Optional[torch.Tensor] = None, mode: Optional[MixedCommMode] = None) -> torch.Tensor:
        return run_mixed_comm(op, self, x_in, x_out, mode)


def _store(res: torch.Tensor, x_out: Optional[torch.Tensor]) -> torch.Tensor:
    if x_out is None:
        return res
    x_out.copy_(res)
    return x_out


def _all_reduce(x: torch.Tensor, group) -> torch.Tensor:
    y = x.clone()
    dist.all_reduce(y, group=group)
    return y


def _all_gather(x: torch.Tensor, group) -> torch.Tensor:
    n = dist.get_world_size(group)
    parts = [torch.empty_like(x) for _ in range(n)]
    dist.all_gather(parts, x.contiguous(), group=group)
    return torch.cat(parts, 0)


def _reduce_scatter(x: torch.Tensor, group) -> torch.Tensor:
    n, r = dist.get_world_size(group), dist.get_rank(group)
    if x.shape[0] % n:
        raise ValueError(f"reduce-scatter: leading dimension {x.shape[0]} is not a multiple of the DP group size {n}")
    if x.is_cuda:
        out = torch.empty(x.shape[0] // n, *x.shape[1:], dtype=x.dtype, device=x.device)
        dist.reduce_scatter_tensor(out, x.contiguous(), group=group)
        return out
    return _all_reduce(x, group).unflatten(0, (n, -1))[r].clone()          # gloo has no reduce-scatter


def run_mixed_comm(op: MixedCommOp, handler: MixedCommHandler, x_in: torch.Tensor, x_out: Optional[torch.Tensor] = None,
                   mode: Optional[MixedCommMode] = None) -> torch.Tensor:
    """Run one collective.  ``x_in`` at least 2-D, of the handler's dtype; ``mode=None`` = autotune when enabled, else NCCL."""
    op = MixedCommOp(op)
    if op not in handler.valid_op_list:
        raise ValueError(f"{op.name} needs {'a TP' if op == MixedCommOp.ALLREDUCE else 'a DP' if op in (MixedCommOp.ALLGATHER, MixedCommOp.REDUCESCATTER) else 'both a TP and a DP'} "
                         f"axis larger than 1 (tp_size {handler.para_info.tp_size}, dp_size {handler.para_info.dp_size})")
    if x_in.dim() < 2:
        raise ValueError("x_in must be at least 2-D")
    if x_in.dtype != handler.dtype:
        raise ValueError(f"x_in dtype {x_in.dtype} does not match the handler's {handler.dtype}")
    if mode is None:
        mode = MixedCommMode.AUTOTUNE if handler.use_autotune else handler.valid_mode_list[-1]
    if mode == MixedCommMode.AUTOTUNE:
        mode = handler.select_autotune_mode(op, x_in)
    if mode not in handler.valid_mode_list:
        raise ValueError(f"mode {MixedCommMode(mode).name} is not available for this topology / device (valid: {[m.name for m in handler.valid_mode_list]})")
    tp, dp = handler.tp_comm_group, handler.dp_comm_group
    if mode in _FUSED:                                          # whole-world group on a CUDA device (checked by valid_mode_list)
        if op == MixedCommOp.ALLREDUCE:
            return _store(handler._ar.all_reduce(x_in), x_out)
        if op == MixedCommOp.ALLGATHER:
            return _store(handler._coll.all_gather(x_in), x_out)
        return _store(handler._coll.reduce_scatter(x_in), x_out)
    if op == MixedCommOp.ALLREDUCE:
        return _store(_all_reduce(x_in, tp), x_out)
    if op == MixedCommOp.ALLGATHER:
        return _store(_all_gather(x_in, dp), x_out)
    if op == MixedCommOp.REDUCESCATTER:
        return _store(_reduce_scatter(x_in, dp), x_out)
    if op == MixedCommOp.ALLREDUCE_ALLGATHER:
        return _store(_all_gather(_all_reduce(x_in, tp), dp), x_out)
    return _store(_all_reduce(_reduce_scatter(x_in, dp), tp), x_out)


get_mixed_comm_module = _jit_acc.module_accessor("comm_collectives")
