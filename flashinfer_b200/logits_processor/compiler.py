"""Fixed-point application of the fusion rules (reference flashinfer/logits_processor/compiler.py)."""
from __future__ import annotations

from typing import List, Optional, Sequence

from .fusion_rules import DEFAULT_RULES, FusionRule
from .op import Op
from .types import CompileError


class Compiler:
    def __init__(self, rules: Optional[Sequence[FusionRule]] = None) -> None:
        self.rules = sorted(list(rules) if rules is not None else list(DEFAULT_RULES), key=lambda r: -r.priority)
        self.validity_checks: list = []

    fusion_rules = property(lambda self: self.rules)

    def register_fusion_rule(self, rule: FusionRule) -> None:
        self.rules.append(rule)
        self.rules.sort(key=lambda r: -r.priority)

    def register_validity_check(self, check) -> None:
        self.validity_checks.append(check)

    def compile(self, ops: List[Op]) -> List[Op]:
        if not ops:
            raise CompileError("Cannot compile empty operator list")
        ops = list(ops)
        for check in self.validity_checks:
            check(ops)
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


def compile_pipeline(ops: List[Op], custom_fusion_rules: Optional[Sequence[FusionRule]] = None,
                     custom_validity_checks: Optional[Sequence] = None) -> List[Op]:
    """Default rules and checks plus the caller's (reference compiler.py :117)."""
    from .validators import get_default_validity_checks

    c = Compiler(list(DEFAULT_RULES))
    for rule in custom_fusion_rules or []:
        c.register_fusion_rule(rule)
    for check in list(get_default_validity_checks()) + list(custom_validity_checks or []):
        c.register_validity_check(check)
    return c.compile(ops)
