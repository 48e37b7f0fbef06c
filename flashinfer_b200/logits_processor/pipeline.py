"""LogitsPipe: processors -> legalised primitive ops -> fused ops, then a plain call chain
(reference flashinfer/logits_processor/pipeline.py).  The pieces live in the sibling modules, like in the reference:
``types`` (tensor kinds, errors), ``op`` (Op / ParameterizedOp), ``operators`` (primitive ops bound to the native sampling
kernels), ``processors`` (Temperature / Softmax / TopK / TopP / MinP / Sample), ``legalization``, ``fusion_rules``, ``compiler``,
``validators``."""
from __future__ import annotations

from typing import Any, List, Optional

import torch

from .compiler import Compiler, compile_pipeline  # noqa: F401
from .fusion_rules import DEFAULT_RULES, FusionRule  # noqa: F401
from .legalization import infer_initial_type, legalize_processors
from .op import Op, ParameterizedOp  # noqa: F401
from .operators import *  # noqa: F401,F403
from .processors import LogitsProcessor, MinP, Sample, Softmax, Temperature, TopK, TopP  # noqa: F401
from .types import CompileError, LegalizationError, TaggedTensor, TensorType  # noqa: F401
from .validators import validate_ops, validate_pipeline


class LogitsPipe:
    def __init__(self, processors: List[LogitsProcessor], compile: bool = True, input_type: Optional[TensorType] = None,
                 custom_fusion_rules: Optional[List[FusionRule]] = None, custom_validity_checks=None) -> None:
        if not processors:
            raise ValueError("Pipeline cannot be empty")
        self.processors = list(processors)
        validate_pipeline(self.processors)
        self._initial_type = input_type or infer_initial_type(self.processors)
        self.ops = legalize_processors(self.processors, self._initial_type)
        validate_ops(self.ops, custom_validity_checks)
        self._rules, self._checks = custom_fusion_rules, custom_validity_checks
        self.compiled_ops: Optional[List[Op]] = None
        if compile:
            self.compile()

    def compile(self, custom_fusion_rules: Optional[List[FusionRule]] = None, custom_validity_checks=None) -> None:
        """Fuse the legalised ops; rules / checks given here are added to the ones given to the constructor (reference pipeline.py :168)."""
        rules = list(self._rules or []) + list(custom_fusion_rules or [])
        checks = list(self._checks or []) + list(custom_validity_checks or [])
        try:
            self.compiled_ops = compile_pipeline(self.ops, rules, checks)
        except CompileError as e:
            raise ValueError(f"Compilation failed: {e}") from e

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
