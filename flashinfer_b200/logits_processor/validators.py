"""Pipeline validity checks (reference flashinfer/logits_processor/validators.py).  Two layers: checks on the processor list before
legalisation (a pipeline samples at most once, and last) and - the reference's form, also what ``custom_validity_checks`` receive -
checks on the legalised op list (``single_softmax_rule``, ``indices_terminal_rule``)."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

from .op import Op
from .operators import SoftmaxOp, TempSoftmaxOp
from .processors import LogitsProcessor, Sample
from .types import CompileError, LegalizationError, TensorType  # noqa: F401  (CompileError is the reference's export of this module)

ValidityCheck = Callable[[List[Op]], None]


def _sample_is_last(processors: Sequence[LogitsProcessor]) -> None:
    for i, p in enumerate(processors):
        if isinstance(p, Sample) and i != len(processors) - 1:
            raise LegalizationError("Sample must be the last processor of a pipeline (it turns probabilities / logits into token ids)")


def _no_duplicate_sample(processors: Sequence[LogitsProcessor]) -> None:
    if sum(isinstance(p, Sample) for p in processors) > 1:
        raise LegalizationError("a pipeline can sample only once")


DEFAULT_VALIDATORS: List[Callable[[Sequence[LogitsProcessor]], None]] = [_no_duplicate_sample, _sample_is_last]


def single_softmax_rule(ops: List[Op]) -> None:
    """At most one softmax (plain or fused with the temperature) per pipeline."""
    if sum(isinstance(o, (SoftmaxOp, TempSoftmaxOp)) for o in ops) > 1:
        raise CompileError("Multiple Softmax operators found. Only one Softmax is allowed per pipeline.")


def indices_terminal_rule(ops: List[Op]) -> None:
    """Nothing may follow an op that produces token ids."""
    for op, nxt in zip(ops, ops[1:]):
        if op.OUT == TensorType.INDICES:
            raise CompileError(f"No operator may follow one that outputs Indices: found {nxt} after {op}")


def get_default_validity_checks() -> List[ValidityCheck]:
    return [single_softmax_rule, indices_terminal_rule]


def validate_pipeline(ops, custom_checks: Optional[Sequence[ValidityCheck]] = None) -> None:
    """The reference's form (validators.py :87) takes the legalised ops: default op-level checks, then ``custom_checks``.  A list of
    high-level processors gets the processor-level checks (raises :class:`LegalizationError` for pipelines that cannot be lowered)."""
    if not ops:
        raise CompileError("Pipeline cannot be empty")
    if all(isinstance(o, Op) for o in ops):
        validate_ops(list(ops), custom_checks)
        return
    for check in DEFAULT_VALIDATORS:
        check(ops)


def validate_ops(ops: List[Op], custom_validity_checks: Optional[Sequence[ValidityCheck]] = None) -> None:
    """Op-level checks on the legalised pipeline: the defaults, then the caller's (they raise whatever they like)."""
    for check in get_default_validity_checks() + list(custom_validity_checks or []):
        check(ops)
