"""Peephole fusion of neighbouring primitive ops onto the fused sampling kernels (reference flashinfer/logits_processor/fusion_rules.py)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple, Union

from .op import Op
from .operators import (MinPSampleOp, SampleLogitsOp, TempSoftmaxOp, TopKSampleOp, TopKTopPSampleOp, TopPSampleOp)


# ------------------------------------------------------------------ fusion
@dataclass
class FusionRule:
    """``pattern``: the window of neighbouring ops to replace - op names (``"softmax"``) or op classes (``SoftmaxOp``, the reference's
    form); ``build(window) -> Op``; ``guard(window) -> bool`` (optional); higher ``priority`` (``prio`` in the reference) is tried first."""
    pattern: Tuple[Union[str, type], ...]
    build: Optional[Callable[[List[Op]], Op]] = None
    guard: Optional[Callable[[List[Op]], bool]] = None
    priority: int = 0
    prio: Optional[int] = None

    def __post_init__(self):
        if self.prio is not None:
            self.priority = int(self.prio)
        if self.build is None:
            raise TypeError("FusionRule needs a build function")

    def matches(self, window: List[Op]) -> bool:
        return len(window) == len(self.pattern) and all(
            (o.name == p) if isinstance(p, str) else isinstance(o, p) for o, p in zip(window, self.pattern)) and (
            self.guard is None or bool(self.guard(window)))


def _merge_static(ops: List[Op]) -> Dict[str, Any]:
    d: Dict[str, Any] = {}
    for o in ops:
        d.update(o.static)
    return d


DEFAULT_RULES: List[FusionRule] = [
    FusionRule(("topk_renorm_probs", "topp_renorm_probs", "sample_probs"), lambda ops: TopKTopPSampleOp(**_merge_static(ops)), priority=3),
    FusionRule(("temperature", "softmax"), lambda ops: TempSoftmaxOp(**_merge_static(ops)), priority=2),
    FusionRule(("topk_renorm_probs", "sample_probs"), lambda ops: TopKSampleOp(**_merge_static(ops)), priority=1),
    FusionRule(("topp_renorm_probs", "sample_probs"), lambda ops: TopPSampleOp(**_merge_static(ops)), priority=1),
    FusionRule(("minp_renorm_probs", "sample_probs"), lambda ops: MinPSampleOp(**_merge_static(ops)), priority=1),
    FusionRule(("softmax", "sample_probs"), lambda ops: SampleLogitsOp(**_merge_static(ops)), priority=1),
]


# ---- the reference's builder / guard names (flashinfer/logits_processor/fusion_rules.py)
def joint_topk_topp_sampleprobs_guard(window: List[Op]) -> bool:
    """Fuse TopK + TopP + Sample into the joint kernel only when TopK was declared with ``joint_topk_topp=True``."""
    return bool(_merge_static(window).get("joint"))


def build_temperature_softmax(window: List[Op]) -> Op:
    return TempSoftmaxOp(**_merge_static(window))


def build_topk_sampling(window: List[Op]) -> Op:
    return TopKSampleOp(**_merge_static(window))


def build_topp_sampling(window: List[Op]) -> Op:
    return TopPSampleOp(**_merge_static(window))


def build_minp_sampling(window: List[Op]) -> Op:
    return MinPSampleOp(**_merge_static(window))


def get_default_fusion_rules() -> List[FusionRule]:
    return list(DEFAULT_RULES)
