"""Processors -> typed primitive ops, with type inference of the pipeline input (reference flashinfer/logits_processor/legalization.py)."""
from __future__ import annotations

from typing import List, Optional, Sequence

from .op import Op
from .processors import LogitsProcessor, MinP, Softmax, Temperature, TopP
from .types import LegalizationError, TensorType


def infer_initial_type(processors: Sequence[LogitsProcessor]) -> TensorType:
    first = processors[0]
    if isinstance(first, (Temperature, Softmax)):
        return TensorType.LOGITS
    if isinstance(first, (TopP, MinP)):
        return TensorType.PROBS
    raise LegalizationError(f"cannot infer the input type from {type(first).__name__}; pass input_type=")


def legalize_processors(processors: Sequence[LogitsProcessor], initial_type: TensorType = TensorType.LOGITS, *,
                        input_type: Optional[TensorType] = None) -> List[Op]:
    ops: List[Op] = []
    cur = input_type if input_type is not None else initial_type
    for p in processors:
        if cur == TensorType.INDICES:
            raise LegalizationError("no processor may follow Sample")
        lowered = p.legalize(cur)
        for op in lowered:
            if op.IN != cur:
                raise LegalizationError(f"{op} does not accept {cur.value}")
            cur = op.OUT
        ops += lowered
    return ops


def validate_processor_chain(processors: Sequence[LogitsProcessor]) -> None:
    """Raises :class:`LegalizationError` when the chain is empty, its input type cannot be inferred, or it cannot be lowered."""
    if not processors:
        raise LegalizationError("Processor chain cannot be empty")
    try:
        legalize_processors(processors, infer_initial_type(processors))
    except LegalizationError:
        raise
    except Exception as exc:  # noqa: BLE001
        raise LegalizationError(f"Processor chain validation failed: {exc}") from exc
