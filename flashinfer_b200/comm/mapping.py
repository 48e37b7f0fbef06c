"""Rank algebra for TP / CP / PP / MoE-TP / MoE-EP / attention-DP layouts.

Parity: reference flashinfer/comm/mapping.py:21-475.  Layout (fastest to slowest): tp, cp, pp —
``rank = pp_rank * tp_size * cp_size + cp_rank * tp_size + tp_rank``; inside a TP group
``tp_rank = moe_tp_rank * moe_ep_size + moe_ep_rank`` (EP groups are contiguous, MoE-TP groups are strided).
"""
from __future__ import annotations

from typing import List, Optional


class Mapping:
    def __init__(self, world_size: int = 1, rank: int = 0, gpus_per_node: int = 8, cp_size: int = 1, cp_config=None,
                 tp_size: int = 1, pp_size: int = 1, moe_cluster_size: int = -1, moe_tp_size: int = -1, moe_ep_size: int = -1,
                 attn_tp_size: int = -1, attn_cp_size: int = -1, auto_parallel: bool = False, enable_attention_dp: bool = False) -> None:
        if moe_cluster_size == -1:
            moe_cluster_size = 1
        if moe_tp_size == -1 and moe_ep_size == -1:
            moe_tp_size, moe_ep_size = tp_size // moe_cluster_size, 1
        elif moe_tp_size == -1:
            moe_tp_size = tp_size // (moe_ep_size * moe_cluster_size)
        elif moe_ep_size == -1:
            moe_ep_size = tp_size // (moe_tp_size * moe_cluster_size)
        if attn_tp_size == -1 and attn_cp_size == -1:
            attn_tp_size, attn_cp_size = tp_size * cp_size, 1
            if cp_size > 1 and cp_config is not None:
                attn_tp_size, attn_cp_size = tp_size, cp_size
        elif attn_tp_size == -1:
            attn_tp_size = cp_size * tp_size // attn_cp_size
        elif attn_cp_size == -1:
            attn_cp_size = cp_size * tp_size // attn_tp_size
        if not auto_parallel and tp_size * pp_size * cp_size != world_size:
            raise ValueError(f"world_size {world_size} != tp {tp_size} * pp {pp_size} * cp {cp_size}")
        if moe_tp_size * moe_ep_size * moe_cluster_size != tp_size:
            raise ValueError(f"tp_size {tp_size} != moe_tp {moe_tp_size} * moe_ep {moe_ep_size} * moe_cluster {moe_cluster_size}")
        if attn_tp_size * attn_cp_size != tp_size * cp_size:
            raise ValueError("attn_tp_size * attn_cp_size must equal tp_size * cp_size")
        if not 0 <= rank < max(world_size, 1):
            raise ValueError(f"rank {rank} out of range for world_size {world_size}")
        self.world_size, self.gpus_per_node = world_size, gpus_per_node
        self.tp_size, self.pp_size, self.cp_size, self.cp_config = tp_size, pp_size, cp_size, cp_config or {}
        self.moe_tp_size, self.moe_ep_size, self.moe_cluster_size = moe_tp_size, moe_ep_size, moe_cluster_size
        self.attn_tp_size, self.attn_cp_size = attn_tp_size, attn_cp_size
        self.auto_parallel, self.enable_attention_dp = auto_parallel, enable_attention_dp
        self.rank = rank

    # ---- rank decomposition
    @property
    def rank(self) -> int:
        return self._rank

    @rank.setter
    def rank(self, r: int) -> None:
        self._rank = r

    @property
    def tp_rank(self) -> int:
        return 0 if self.auto_parallel else self._rank % self.tp_size

    @property
    def cp_rank(self) -> int:
        return 0 if self.auto_parallel else (self._rank // self.tp_size) % self.cp_size

    @property
    def pp_rank(self) -> int:
        return 0 if self.auto_parallel else self._rank // (self.tp_size * self.cp_size)

    @property
    def moe_ep_rank(self) -> int:
        return self.tp_rank % self.moe_ep_size

    @property
    def moe_cluster_rank(self) -> int:
        return (self.tp_rank // self.moe_ep_size) % self.moe_cluster_size

    @property
    def moe_tp_rank(self) -> int:
        return self.tp_rank // (self.moe_ep_size * self.moe_cluster_size)

    @property
    def attn_tp_rank(self) -> int:
        return (self._rank % (self.tp_size * self.cp_size)) % self.attn_tp_size

    @property
    def attn_cp_rank(self) -> int:
        return (self._rank % (self.tp_size * self.cp_size)) // self.attn_tp_size

    @property
    def node_rank(self) -> int:
        return self._rank // self.gpus_per_node

    @property
    def local_rank(self) -> int:
        return self._rank % self.gpus_per_node

    # ---- groups (lists of global ranks)
    @property
    def tp_group(self) -> List[int]:
        base = self._rank - self.tp_rank
        return [base + i for i in range(self.tp_size)]

    @property
    def cp_group(self) -> List[int]:
        base = self.pp_rank * self.tp_size * self.cp_size + self.tp_rank
        return [base + i * self.tp_size for i in range(self.cp_size)]

    @property
    def pp_group(self) -> List[int]:
        stride = self.tp_size * self.cp_size
        base = self._rank % stride
        return [base + i * stride for i in range(self.pp_size)]

    @property
    def moe_ep_group(self) -> List[int]:
        base = self._rank - self.tp_rank + (self.tp_rank // self.moe_ep_size) * self.moe_ep_size
        return [base + i for i in range(self.moe_ep_size)]

    @property
    def moe_tp_group(self) -> List[int]:
        stride = self.moe_ep_size * self.moe_cluster_size
        base = self._rank - self.tp_rank + self.tp_rank % stride
        return [base + i * stride for i in range(self.moe_tp_size)]

    @property
    def moe_cluster_group(self) -> List[int]:
        base = self._rank - self.tp_rank + self.moe_tp_rank * self.moe_ep_size * self.moe_cluster_size + self.moe_ep_rank
        return [base + i * self.moe_ep_size for i in range(self.moe_cluster_size)]

    # ---- predicates / helpers
    def has_tp(self) -> bool:
        return self.tp_size > 1

    def has_cp(self) -> bool:
        return self.cp_size > 1

    def has_pp(self) -> bool:
        return self.pp_size > 1

    def has_moe_tp(self) -> bool:
        return self.moe_tp_size > 1

    def has_moe_ep(self) -> bool:
        return self.moe_ep_size > 1

    def has_moe_cluster(self) -> bool:
        return self.moe_cluster_size > 1

    def is_first_pp_rank(self) -> bool:
        return self.pp_rank == 0

    def is_last_pp_rank(self) -> bool:
        return self.pp_rank == self.pp_size - 1

    def is_multi_node(self) -> bool:
        return self.world_size > self.gpus_per_node

    def prev_pp_rank(self) -> int:
        stride = self.tp_size * self.cp_size
        return (self._rank - stride) % self.world_size

    def next_pp_rank(self) -> int:
        stride = self.tp_size * self.cp_size
        return (self._rank + stride) % self.world_size

    def prev_cp_rank(self) -> int:
        g = self.cp_group
        return g[(self.cp_rank - 1) % self.cp_size]

    def next_cp_rank(self) -> int:
        g = self.cp_group
        return g[(self.cp_rank + 1) % self.cp_size]

    def pp_layers(self, num_layers: int, layer_fusion=None) -> List[int]:
        """Layers owned by this pipeline stage (remainder layers go to the first stages)."""
        base, rem = divmod(num_layers, self.pp_size)
        start = self.pp_rank * base + min(self.pp_rank, rem)
        return list(range(start, start + base + (1 if self.pp_rank < rem else 0)))

    def ep_experts(self, num_experts: int) -> List[int]:
        base, rem = divmod(num_experts, self.moe_ep_size)
        start = self.moe_ep_rank * base + min(self.moe_ep_rank, rem)
        return list(range(start, start + base + (1 if self.moe_ep_rank < rem else 0)))

    def get_node_rank(self, rank: int) -> int:
        return rank // self.gpus_per_node

    def get_local_rank(self, rank: int) -> int:
        return rank % self.gpus_per_node

    def to_dict(self) -> dict:
        return {k: getattr(self, k) for k in ("world_size", "rank", "gpus_per_node", "cp_size", "tp_size", "pp_size",
                                              "moe_tp_size", "moe_ep_size", "moe_cluster_size", "attn_tp_size", "attn_cp_size",
                                              "enable_attention_dp")}

    @classmethod
    def from_dict(cls, d: dict) -> "Mapping":
        return cls(**d)

    def __eq__(self, other) -> bool:
        return isinstance(other, Mapping) and self.to_dict() == other.to_dict()

    def __hash__(self) -> int:
        return hash(tuple(sorted(self.to_dict().items())))

    def __repr__(self) -> str:
        return f"Mapping({self.to_dict()})"
