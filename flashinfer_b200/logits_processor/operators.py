"""The primitive ops a pipeline is legalised into, each bound to a native kernel of flashinfer_b200.sampling
(reference flashinfer/logits_processor/operators.py)."""
from __future__ import annotations

import torch

from .. import sampling as S
from .op import ParameterizedOp
from .types import TensorType


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

# the reference's class names for the same ops (flashinfer/logits_processor/operators.py), so user fusion rules / validity checks written
# against ``isinstance(op, ProbsTopKOp)`` work unchanged
LogitsTopKOp, ProbsTopKOp, TopPOp, MinPOp = TopKLogitsOp, TopKProbsOp, TopPProbsOp, MinPProbsOp
ProbsSampleOp, LogitsSampleOp = SampleProbsOp, SampleLogitsOp
FusedTemperatureSoftmaxOp, FusedProbsTopKSampleOp, FusedProbsTopPSampleOp = TempSoftmaxOp, TopKSampleOp, TopPSampleOp
FusedProbsMinPSampleOp, FusedProbsTopKTopPSampleOp = MinPSampleOp, TopKTopPSampleOp

__all__ = [n for n in dir() if n.endswith("Op")]
