"""Declarative logits post-processing pipelines.

Parity: reference flashinfer/logits_processor/ (LogitsPipe, processors Temperature / Softmax / TopK / TopP / MinP /
Sample, legalisation to typed ops, fusion rules, compile_pipeline).

The pipeline is legalised into typed primitive ops (LOGITS -> LOGITS, LOGITS -> PROBS, PROBS -> PROBS, PROBS -> INDICES)
and then peephole-fused onto the native fused kernels of :mod:`flashinfer_b200.sampling`:
  Temperature + Softmax           -> softmax(logits, temperature)                 (one kernel)
  TopK + Sample (probs)           -> top_k_sampling_from_probs                    (rejection sampling, no renorm pass)
  TopP + Sample                   -> top_p_sampling_from_probs
  MinP + Sample                   -> min_p_sampling_from_probs
  TopK + TopP + Sample            -> top_k_top_p_sampling_from_probs
  Softmax + Sample                -> sampling_from_logits (Gumbel-max, no probs materialised)
"""
from .pipeline import (  # noqa: F401
    CompileError,
    Compiler,
    FusionRule,
    LegalizationError,
    LogitsPipe,
    LogitsProcessor,
    MinP,
    Op,
    ParameterizedOp,
    Sample,
    Softmax,
    TaggedTensor,
    Temperature,
    TensorType,
    TopK,
    TopP,
    compile_pipeline,
    legalize_processors,
)

from .. import _alias  # noqa: E402

from . import compiler, fusion_rules, legalization, op, operators, processors, types, validators  # noqa: E402,F401  (the reference's module layout)
