"""Peephole fusion of neighbouring primitive ops onto the fused sampling kernels (reference flashinfer/logits_processor/fusion_rules.py)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

from .op import Op
from .operators import (MinPSampleOp, SampleLogitsOp, TempSoftmaxOp, TopKSampleOp, TopKTopPSampleOp, TopPSampleOp)


# ------------------------------------------------------------------ fusion
@dataclass
class FusionRule:
    pattern: Tuple[str, ...]
    build: Callable[[List[Op]], Op]
    guard: Optional[Callable[[List[Op]], bool]] = None
    priority: int = 0


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
