"""Tensor kinds flowing through a logits pipeline and the pipeline errors (reference flashinfer/logits_processor/types.py)."""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum

import torch


class TensorType(Enum):
    LOGITS = "logits"
    PROBS = "probs"
    INDICES = "indices"


@dataclass
class TaggedTensor:
    data: torch.Tensor
    type: TensorType

    @staticmethod
    def logits(t: torch.Tensor) -> "TaggedTensor":
        return TaggedTensor(t, TensorType.LOGITS)

    @staticmethod
    def probs(t: torch.Tensor) -> "TaggedTensor":
        return TaggedTensor(t, TensorType.PROBS)

    @staticmethod
    def indices(t: torch.Tensor) -> "TaggedTensor":
        return TaggedTensor(t, TensorType.INDICES)

    shape = property(lambda self: self.data.shape)
    device = property(lambda self: self.data.device)
    dtype = property(lambda self: self.data.dtype)

    def size(self, dim=None):
        return self.data.size() if dim is None else self.data.size(dim)


class LegalizationError(ValueError):
    pass


class CompileError(ValueError):
    pass
