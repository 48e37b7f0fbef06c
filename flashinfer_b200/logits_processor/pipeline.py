from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch

from .. import sampling as S


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


class LegalizationError(ValueError):
    pass


class CompileError(ValueError):
    pass


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


def _make_op(name, tin, tout, fn, params=()):
    cls = type(name, (ParameterizedOp,), {"name": name, "IN": tin, "OUT": tout, "params": tuple(params),
                                          "__call__": lambda self, x, **kw: fn(self, x, **kw)})
    return cls


TemperatureOp = _make_op("temperature", TensorType.LOGITS, TensorType.LOGITS,
                         lambda self, x, **kw: x / (self._get(kw, "temperature") if not isinstance(self._get(kw, "temperature"), torch.Tensor)
                                                    else self._get(kw, "temperature").reshape(-1, 1)), ("temperature",))
SoftmaxOp = _make_op("softmax", TensorType.LOGITS, TensorType.PROBS, lambda self, x, **kw: S.softmax(x))
TempSoftmaxOp = _make_op("temperature_softmax", TensorType.LOGITS, TensorType.PROBS,
                         lambda self, x, **kw: S.softmax(x, self._get(kw, "temperature")), ("temperature",))
TopKLogitsOp = _make_op("topk_mask_logits", TensorType.LOGITS, TensorType.LOGITS,
                        lambda self, x, **kw: S.top_k_mask_logits(x, self._get(kw, "top_k")), ("top_k",))
TopKProbsOp = _make_op("topk_renorm_probs", TensorType.PROBS, TensorType.PROBS,
                       lambda self, x, **kw: S.top_k_renorm_probs(x, self._get(kw, "top_k")), ("top_k",))
TopPProbsOp = _make_op("topp_renorm_probs", TensorType.PROBS, TensorType.PROBS,
                       lambda self, x, **kw: S.top_p_renorm_probs(x, self._get(kw, "top_p")), ("top_p",))


def _minp_renorm(self, x, **kw):
    p = self._get(kw, "min_p")
    p = p.reshape(-1, 1) if isinstance(p, torch.Tensor) else p
    keep = x >= x.amax(-1, keepdim=True) * p
    y = torch.where(keep, x, torch.zeros_like(x))
    return y / y.sum(-1, keepdim=True)


MinPProbsOp = _make_op("minp_renorm_probs", TensorType.PROBS, TensorType.PROBS, _minp_renorm, ("min_p",))


def _gen(kw):
    return dict(indices=kw.get("indices"), generator=kw.get("generator"))


SampleProbsOp = _make_op("sample_probs", TensorType.PROBS, TensorType.INDICES,
                         lambda self, x, **kw: S.sampling_from_probs(x, deterministic=self.static.get("deterministic", True), **_gen(kw)))
SampleLogitsOp = _make_op("sample_logits", TensorType.LOGITS, TensorType.INDICES,
                          lambda self, x, **kw: S.sampling_from_logits(x, deterministic=self.static.get("deterministic", True), **_gen(kw)))
TopKSampleOp = _make_op("topk_sample", TensorType.PROBS, TensorType.INDICES,
                        lambda self, x, **kw: S.top_k_sampling_from_probs(x, self._get(kw, "top_k"), **_gen(kw)), ("top_k",))
TopPSampleOp = _make_op("topp_sample", TensorType.PROBS, TensorType.INDICES,
                        lambda self, x, **kw: S.top_p_sampling_from_probs(x, self._get(kw, "top_p"), **_gen(kw)), ("top_p",))
MinPSampleOp = _make_op("minp_sample", TensorType.PROBS, TensorType.INDICES,
                        lambda self, x, **kw: S.min_p_sampling_from_probs(x, self._get(kw, "min_p"), **_gen(kw)), ("min_p",))
TopKTopPSampleOp = _make_op("topk_topp_sample", TensorType.PROBS, TensorType.INDICES,
                            lambda self, x, **kw: S.top_k_top_p_sampling_from_probs(
                                x, self._get(kw, "top_k"), self._get(kw, "top_p"),
                                filter_apply_order="joint" if self.static.get("joint") else "top_k_first", **_gen(kw)),
                            ("top_k", "top_p"))


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


def infer_initial_type(processors: Sequence[LogitsProcessor]) -> TensorType:
    first = processors[0]
    if isinstance(first, (Temperature, Softmax)):
        return TensorType.LOGITS
    if isinstance(first, (TopP, MinP)):
        return TensorType.PROBS
    raise LegalizationError(f"cannot infer the input type from {type(first).__name__}; pass input_type=")


def legalize_processors(processors: Sequence[LogitsProcessor], input_type: TensorType) -> List[Op]:
    ops: List[Op] = []
    cur = input_type
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
                    if tuple(o.name for o in window) == rule.pattern and (rule.guard is None or rule.guard(window)):
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


class LogitsPipe:
    def __init__(self, processors: List[LogitsProcessor], compile: bool = True, input_type: Optional[TensorType] = None,
                 custom_fusion_rules: Optional[List[FusionRule]] = None, custom_validity_checks=None) -> None:
        if not processors:
            raise ValueError("Pipeline cannot be empty")
        self.processors = list(processors)
        self._initial_type = input_type or infer_initial_type(self.processors)
        self.ops = legalize_processors(self.processors, self._initial_type)
        self._rules = custom_fusion_rules
        self.compiled_ops: Optional[List[Op]] = None
        if compile:
            self.compile()

    def compile(self) -> None:
        self.compiled_ops = compile_pipeline(self.ops, self._rules)

    @property
    def initial_type(self) -> TensorType:
        return self._initial_type

    def __call__(self, x, **kwargs: Any) -> torch.Tensor:
        if isinstance(x, TaggedTensor):
            if x.type != self._initial_type:
                raise ValueError(f"pipeline expects {self._initial_type.value}, got {x.type.value}")
            x = x.data
        for op in (self.compiled_ops if self.compiled_ops is not None else self.ops):
            x = op(x, **kwargs)
        return x

    def __repr__(self) -> str:
        return "LogitsPipe(" + " -> ".join(repr(o) for o in (self.compiled_ops or self.ops)) + ")"
