"""Typed primitive ops (reference flashinfer/logits_processor/op.py)."""
from __future__ import annotations

from typing import Any, Dict, Tuple

import torch

from .types import TensorType


class Op:
    """Typed primitive: ``IN`` -> ``OUT`` with a name used by the fusion rules."""
    name: str = "op"
    IN: TensorType = TensorType.LOGITS
    OUT: TensorType = TensorType.LOGITS

    def __init__(self, **static: Any) -> None:
        self.static = static

    def __call__(self, x: torch.Tensor, **params: Any) -> torch.Tensor:
        raise NotImplementedError

    def __repr__(self) -> str:
        return f"{self.name}({self.IN.value}->{self.OUT.value})"


class ParameterizedOp(Op):
    """Op whose run-time parameters (``temperature``, ``top_k`` ...) are looked up in the call kwargs."""
    params: Tuple[str, ...] = ()

    def _get(self, params: Dict[str, Any], key: str):
        if key in self.static and self.static[key] is not None:
            return self.static[key]
        if key not in params:
            raise ValueError(f"{self.name}: missing run-time parameter '{key}'")
        return params[key]
