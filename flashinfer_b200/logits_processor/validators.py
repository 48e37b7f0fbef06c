"""Pipeline validity checks run before legalisation (reference flashinfer/logits_processor/validators.py)."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

from .processors import LogitsProcessor, Sample
from .types import LegalizationError


def _sample_is_last(processors: Sequence[LogitsProcessor]) -> None:
    for i, p in enumerate(processors):
        if isinstance(p, Sample) and i != len(processors) - 1:
            raise LegalizationError("Sample must be the last processor of a pipeline (it turns probabilities / logits into token ids)")


def _no_duplicate_sample(processors: Sequence[LogitsProcessor]) -> None:
    if sum(isinstance(p, Sample) for p in processors) > 1:
        raise LegalizationError("a pipeline can sample only once")


DEFAULT_VALIDATORS: List[Callable[[Sequence[LogitsProcessor]], None]] = [_no_duplicate_sample, _sample_is_last]


def validate_pipeline(processors: Sequence[LogitsProcessor], custom_validity_checks: Optional[Sequence[Callable]] = None) -> None:
    """Raises :class:`LegalizationError` (or whatever a custom check raises) for pipelines that cannot be lowered."""
    for check in list(DEFAULT_VALIDATORS) + list(custom_validity_checks or []):
        check(processors)
