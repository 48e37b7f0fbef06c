"""Attention variants as Python objects (reference flashinfer/cute_dsl/attention/fusion/variant.py).

In the reference a variant is a class whose ``@cute.jit`` methods are traced into the DSL kernel.  Here a variant lowers to one of
two things the tcgen05 prefill kernel (csrc/attention/prefill_sm100.cu) already knows:

* a *built-in* feature selected at plan() / run() time - ALiBi slopes, logits soft-cap, attention sinks;
* a *compiled hook*: a C++ ``LogitsTransform`` / ``LogitsMask`` declaration handed to
  ``jit.gen_customize_batch_prefill_module``, which builds a private copy of the kernel with the hook inlined into its softmax
  pass (RPE below; user subclasses give ``cuda_decl`` + the tensors / scalars it reads).

Positions: ``kv_idx`` counts keys from the start of the request, the query position is ``kv_len - qo_len + qo_idx`` (queries are the
last tokens).  Hooks see ``logits = q.k * sm_scale``; additive biases are added after that scaling."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch


class AttentionVariant:
    """Base class.  Subclasses either set the built-in fields or describe a compiled hook:

    ``cuda_decl``   C++ source of ``struct <name> : VariantDefaults { ... }`` overriding ``LogitsTransform`` and / or ``LogitsMask``
    ``tensors``     ordered ``{name: tensor}`` the hook reads (visible under their names as typed const pointers)
    ``scalars``     ordered ``{name: float}`` (visible as floats)"""

    name: str = "StandardAttention"
    cuda_decl: str = ""
    alibi_slopes: Optional[torch.Tensor] = None
    logits_soft_cap: float = 0.0
    sinks: Optional[torch.Tensor] = None

    @property
    def tensors(self) -> Dict[str, torch.Tensor]:
        return {}

    @property
    def scalars(self) -> Dict[str, float]:
        return {}

    @property
    def extra_params(self):
        ts = list(self.tensors.values())
        return ts[0] if len(ts) == 1 else (ts or None)

    @property
    def is_compiled_hook(self) -> bool:
        return bool(self.cuda_decl)

    def jit_args(self, uri: str, q_dtype, kv_dtype, head_dim_qk: int, head_dim_vo: int) -> List:
        """Positional arguments of ``jit.gen_customize_batch_prefill_module`` after ``backend``."""
        ctype = {torch.float32: "float", torch.float16: "half", torch.bfloat16: "bfloat16", torch.int32: "int32_t", torch.int64: "int64_t",
                 torch.uint8: "uint8_t", torch.bool: "bool"}
        t = self.tensors
        return [uri, q_dtype, kv_dtype, q_dtype, torch.int32, head_dim_qk, head_dim_vo, list(t), [ctype[v.dtype] for v in t.values()],
                list(self.scalars), ["double"] * len(self.scalars), self.name, self.cuda_decl]

    def run_args(self) -> Tuple:
        """Extra positional arguments of the wrapper's ``run()`` (tensors first, then scalars), in declaration order."""
        return tuple(self.tensors.values()) + tuple(float(v) for v in self.scalars.values())


class StandardAttention(AttentionVariant):
    """Plain softmax attention."""


class AttentionWithSink(AttentionVariant):
    """One virtual key per head whose logit is ``sink[head]``: it joins the softmax denominator and contributes no value
    (``exp(sink)`` is folded into the running sum inside the kernel's epilogue)."""

    name = "AttentionWithSink"

    def __init__(self, sink: torch.Tensor):
        self.sinks = sink


class ALiBiAttention(AttentionVariant):
    """``logits + slope[head] * (kv_pos - q_pos)``: the kernel's built-in ALiBi pass with caller-provided slopes."""

    name = "ALiBiAttention"

    def __init__(self, alibi_slopes: torch.Tensor):
        self.alibi_slopes = alibi_slopes

    @staticmethod
    def get_slopes(num_heads: int) -> torch.Tensor:
        """Geometric slope schedule of the ALiBi paper: ``2^(-8 i / n)`` for a power-of-two head count ``n``; other counts take the
        schedule of the power of two below and fill up with every second slope of the next one."""
        def pow2(n: int) -> List[float]:
            base = 2.0 ** (-8.0 / n)
            return [base ** (i + 1) for i in range(n)]

        lo = 1 << (num_heads.bit_length() - 1)
        slopes = pow2(lo)
        if lo != num_heads:
            slopes += pow2(2 * lo)[0::2][: num_heads - lo]
        return torch.tensor(slopes, dtype=torch.float32)


class SoftCappingAttention(AttentionVariant):
    """``cap * tanh(logits / cap)`` (Gemma-2): the kernel's built-in soft-cap."""

    name = "SoftCappingAttention"

    def __init__(self, cap: float = 50.0):
        self.cap = float(cap)
        self.logits_soft_cap = float(cap)


class RPEAttention(AttentionVariant):
    """Learned relative-position bias: ``logits + rpe_table[head, clamp(kv_pos - q_pos + max_rel_dist, 0, 2 * max_rel_dist)]``,
    compiled into the softmax pass as a LogitsTransform hook."""

    name = "RPEAttention"
    cuda_decl = """
struct RPEAttention : VariantDefaults {
  static __device__ __forceinline__ float LogitsTransform(const VariantCtx& ctx, float logits, int kv_idx) {
    const int span = int(rpe_max_rel_dist);
    int rel = kv_idx - (ctx.kv_len - ctx.qo_len + ctx.qo_idx) + span;
    rel = rel < 0 ? 0 : (rel > 2 * span ? 2 * span : rel);
    return logits + rpe_table[ctx.qo_head_idx * (2 * span + 1) + rel];
  }
};
"""

    def __init__(self, rpe_table: torch.Tensor, max_rel_dist: int):
        if rpe_table.dim() != 2 or rpe_table.shape[1] != 2 * max_rel_dist + 1:
            raise ValueError("rpe_table must be [num_qo_heads, 2 * max_rel_dist + 1]")
        self._table = rpe_table.float().contiguous()
        self._max_rel_dist = int(max_rel_dist)

    @property
    def tensors(self):
        return {"rpe_table": self._table}

    @property
    def scalars(self):
        return {"rpe_max_rel_dist": float(self._max_rel_dist)}


class SigmoidAttention(AttentionVariant):
    """Element-wise ``sigmoid(scale * logits + bias)`` instead of softmax.  The tcgen05 prefill kernel normalises rows with an online
    softmax whose rescaling cannot be switched off by a logits hook, so this variant is served by the plain PyTorch path of
    :func:`sigmoid_attention_reference` (small problems / tests); a dedicated epilogue is not written."""

    name = "SigmoidAttention"

    def __init__(self, scale: float = 1.0, bias: float = 0.0):
        self.scale, self.bias = float(scale), float(bias)


class SigmoidTanhAttention(SigmoidAttention):
    """Same function computed as ``0.5 + 0.5 * tanh(x / 2)`` in the reference (one MUFU op); identical result here."""

    name = "SigmoidTanhAttention"


def sigmoid_attention_reference(q, k, v, qo_indptr, kv_indptr, causal: bool, sm_scale: float, variant: SigmoidAttention) -> torch.Tensor:
    """out = sigmoid(scale * (q.k * sm_scale) + bias) @ v per request (no row normalisation)."""
    hq, hkv = q.shape[1], k.shape[1]
    out = torch.zeros(q.shape[0], hq, v.shape[-1], dtype=torch.float32, device=q.device)
    for i in range(qo_indptr.numel() - 1):
        qs, qe, ks, ke = int(qo_indptr[i]), int(qo_indptr[i + 1]), int(kv_indptr[i]), int(kv_indptr[i + 1])
        kf = k[ks:ke].float().repeat_interleave(hq // hkv, 1)
        vf = v[ks:ke].float().repeat_interleave(hq // hkv, 1)
        p = torch.sigmoid(torch.einsum("qhd,khd->hqk", q[qs:qe].float(), kf) * sm_scale * variant.scale + variant.bias)
        if causal:
            qpos = torch.arange(qe - qs, device=q.device)[:, None] + ((ke - ks) - (qe - qs))
            p = p.masked_fill(torch.arange(ke - ks, device=q.device)[None, :] > qpos, 0.0)
        out[qs:qe] = torch.einsum("hqk,khd->qhd", p, vf)
    return out.to(q.dtype)


def tanh_approx(x):
    """Host-side stand-in of the reference's MUFU tanh helper (device code uses ``tanh.approx.f32`` inside the kernel's soft-cap)."""
    return torch.tanh(x) if isinstance(x, torch.Tensor) else math.tanh(x)
