"""User-facing processors: what a caller lists in LogitsPipe([...]) (reference flashinfer/logits_processor/processors.py)."""
from __future__ import annotations

from typing import Any, List, Optional

from .op import Op
from .operators import (MinPProbsOp, SampleLogitsOp, SampleProbsOp, SoftmaxOp, TemperatureOp, TopKLogitsOp, TopKProbsOp, TopPProbsOp)
from .types import LegalizationError, TensorType


# ------------------------------------------------------------------ user-facing processors
class LogitsProcessor:
    """High-level pipeline stage; ``legalize(input_type)`` lowers it to typed ops."""

    def __init__(self, **params: Any) -> None:
        self.params = params

    def legalize(self, input_type: TensorType) -> List[Op]:
        raise NotImplementedError

    def __repr__(self) -> str:
        return f"{type(self).__name__}({self.params})"


class Temperature(LogitsProcessor):
    def legalize(self, input_type):
        if input_type != TensorType.LOGITS:
            raise LegalizationError("Temperature can only be applied to logits")
        return [TemperatureOp(**self.params)]


class Softmax(LogitsProcessor):
    def __init__(self, enable_pdl: Optional[bool] = None, **params):
        super().__init__(**params)

    def legalize(self, input_type):
        if input_type != TensorType.LOGITS:
            raise LegalizationError("Softmax can only be applied to logits")
        return [SoftmaxOp(**self.params)]


class TopK(LogitsProcessor):
    def __init__(self, joint_topk_topp: bool = False, **params):
        super().__init__(**params)
        self.joint = joint_topk_topp

    def legalize(self, input_type):
        if input_type == TensorType.LOGITS:
            return [TopKLogitsOp(**self.params)]
        if input_type == TensorType.PROBS:
            return [TopKProbsOp(joint=self.joint, **self.params)]
        raise LegalizationError("TopK needs logits or probs")


class TopP(LogitsProcessor):
    def legalize(self, input_type):
        if input_type != TensorType.PROBS:
            raise LegalizationError("TopP can only be applied to probabilities (insert Softmax first)")
        return [TopPProbsOp(**self.params)]


class MinP(LogitsProcessor):
    def legalize(self, input_type):
        if input_type != TensorType.PROBS:
            raise LegalizationError("MinP can only be applied to probabilities (insert Softmax first)")
        return [MinPProbsOp(**self.params)]


class Sample(LogitsProcessor):
    def __init__(self, deterministic: bool = True, **params):
        super().__init__(**params)
        self.deterministic = deterministic

    def legalize(self, input_type):
        if input_type == TensorType.PROBS:
            return [SampleProbsOp(deterministic=self.deterministic)]
        if input_type == TensorType.LOGITS:
            return [SampleLogitsOp(deterministic=self.deterministic)]
        raise LegalizationError("Sample needs logits or probs")
