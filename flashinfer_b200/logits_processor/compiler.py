"""Fixed-point application of the fusion rules (reference flashinfer/logits_processor/compiler.py)."""
from __future__ import annotations

from typing import List, Optional, Sequence

from .fusion_rules import DEFAULT_RULES, FusionRule
from .op import Op
from .types import CompileError


class Compiler:
    def __init__(self, rules: Optional[Sequence[FusionRule]] = None) -> None:
        self.rules = sorted(list(rules) if rules is not None else list(DEFAULT_RULES), key=lambda r: -r.priority)

    def compile(self, ops: List[Op]) -> List[Op]:
        ops = list(ops)
        changed = True
        while changed:
            changed = False
            for rule in self.rules:
                n = len(rule.pattern)
                for i in range(len(ops) - n + 1):
                    window = ops[i:i + n]
                    if rule.matches(window):
                        ops[i:i + n] = [rule.build(window)]
                        changed = True
                        break
                if changed:
                    break
        for a, b in zip(ops, ops[1:]):
            if a.OUT != b.IN:
                raise CompileError(f"type mismatch after fusion: {a} -> {b}")
        return ops


def compile_pipeline(ops: List[Op], custom_fusion_rules: Optional[Sequence[FusionRule]] = None) -> List[Op]:
    rules = list(DEFAULT_RULES) + list(custom_fusion_rules or [])
    return Compiler(rules).compile(ops)
