"""``torch.library`` registration of the functional ops: ``torch.ops.flashinfer_b200.*`` custom ops with fake (meta) kernels, so
that ``torch.compile(fullgraph=True)`` / ``torch.export`` trace through code that calls this library without graph breaks.

Parity: reference ``register_custom_op`` / ``register_fake_op`` (flashinfer/utils.py:325-376) and the torch.compile test mode of
its suite (tests/conftest.py:68-133).  The reference turns its registration into a no-op because ``torch.library.custom_op`` adds
dispatch overhead to every eager call; here the eager public API stays the bare ctypes launcher (zero overhead) and the custom
ops live in this opt-in module: ``import flashinfer_b200.torch_ops`` registers them, and ``compiled(fn)`` callers use
``torch.ops.flashinfer_b200.<op>`` (or the thin python aliases below) inside compiled regions.  Mutating ops declare
``mutates_args`` so functionalisation keeps them ordered; fake kernels only allocate outputs of the right shape / dtype / device.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Sequence, Union

import torch

from . import activation as _act
from . import cascade as _cascade
from . import norm as _norm
from . import rope as _rope
from . import sampling as _sampling
from .gemm import dense as _dense

NAMESPACE = "flashinfer_b200"


def register_custom_op(name: str, fn: Optional[Callable] = None, /, *, mutates_args: Union[str, Iterable[str]],
                       device_types: Optional[Union[str, Sequence[str]]] = None, schema: Optional[str] = None) -> Callable:
    """``torch.library.custom_op`` with the reference's signature (flashinfer/utils.py:330)."""
    return torch.library.custom_op(name, fn, mutates_args=mutates_args, device_types=device_types, schema=schema)


def register_fake_op(name: str, fn: Optional[Callable] = None) -> Callable:
    """``torch.library.register_fake`` with the reference's signature (flashinfer/utils.py:365)."""
    return torch.library.register_fake(name, fn)


# ------------------------------------------------------------------ norms / activations (out-of-place results)
@register_custom_op(f"{NAMESPACE}::rmsnorm", mutates_args=())
def rmsnorm(input: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    return _norm.rmsnorm(input, weight, eps)


@register_fake_op(f"{NAMESPACE}::rmsnorm")
def _(input, weight, eps=1e-6):
    return torch.empty_like(input)


@register_custom_op(f"{NAMESPACE}::gemma_rmsnorm", mutates_args=())
def gemma_rmsnorm(input: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    return _norm.gemma_rmsnorm(input, weight, eps)


@register_fake_op(f"{NAMESPACE}::gemma_rmsnorm")
def _(input, weight, eps=1e-6):
    return torch.empty_like(input)


@register_custom_op(f"{NAMESPACE}::fused_add_rmsnorm", mutates_args=("input", "residual"))
def fused_add_rmsnorm(input: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6) -> None:
    _norm.fused_add_rmsnorm(input, residual, weight, eps)


@register_fake_op(f"{NAMESPACE}::fused_add_rmsnorm")
def _(input, residual, weight, eps=1e-6):
    return None


@register_custom_op(f"{NAMESPACE}::silu_and_mul", mutates_args=())
def silu_and_mul(input: torch.Tensor) -> torch.Tensor:
    return _act.silu_and_mul(input)


@register_custom_op(f"{NAMESPACE}::gelu_and_mul", mutates_args=())
def gelu_and_mul(input: torch.Tensor) -> torch.Tensor:
    return _act.gelu_and_mul(input)


@register_custom_op(f"{NAMESPACE}::gelu_tanh_and_mul", mutates_args=())
def gelu_tanh_and_mul(input: torch.Tensor) -> torch.Tensor:
    return _act.gelu_tanh_and_mul(input)


def _half_last(input):
    return input.new_empty(*input.shape[:-1], input.shape[-1] // 2)


for _n in ("silu_and_mul", "gelu_and_mul", "gelu_tanh_and_mul"):
    register_fake_op(f"{NAMESPACE}::{_n}")(_half_last)


# ------------------------------------------------------------------ GEMM
@register_custom_op(f"{NAMESPACE}::linear", mutates_args=())
def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    return _dense.linear(x, weight, bias)


@register_fake_op(f"{NAMESPACE}::linear")
def _(x, weight, bias=None):
    return x.new_empty(*x.shape[:-1], weight.shape[0])


@register_custom_op(f"{NAMESPACE}::linear_gated_silu", mutates_args=())
def linear_gated_silu(x: torch.Tensor, weight_interleaved: torch.Tensor) -> torch.Tensor:
    return _dense.linear_gated_silu(x, weight_interleaved)


@register_fake_op(f"{NAMESPACE}::linear_gated_silu")
def _(x, weight_interleaved):
    return x.new_empty(*x.shape[:-1], weight_interleaved.shape[0] // 2)


# ------------------------------------------------------------------ RoPE (in place)
@register_custom_op(f"{NAMESPACE}::apply_rope_pos_ids_inplace", mutates_args=("q", "k"))
def apply_rope_pos_ids_inplace(q: torch.Tensor, k: torch.Tensor, pos_ids: torch.Tensor, rotary_dim: Optional[int] = None,
                               interleave: bool = False, rope_scale: float = 1.0, rope_theta: float = 1e4) -> None:
    _rope.apply_rope_pos_ids_inplace(q, k, pos_ids, rotary_dim=rotary_dim, interleave=interleave, rope_scale=rope_scale,
                                     rope_theta=rope_theta)


@register_fake_op(f"{NAMESPACE}::apply_rope_pos_ids_inplace")
def _(q, k, pos_ids, rotary_dim=None, interleave=False, rope_scale=1.0, rope_theta=1e4):
    return None


@register_custom_op(f"{NAMESPACE}::apply_llama31_rope_pos_ids_inplace", mutates_args=("q", "k"))
def apply_llama31_rope_pos_ids_inplace(q: torch.Tensor, k: torch.Tensor, pos_ids: torch.Tensor, rotary_dim: Optional[int] = None,
                                       interleave: bool = False, rope_scale: float = 8.0, rope_theta: float = 5e5,
                                       low_freq_factor: float = 1.0, high_freq_factor: float = 4.0,
                                       old_context_len: int = 8192) -> None:
    _rope.apply_llama31_rope_pos_ids_inplace(q, k, pos_ids, rotary_dim=rotary_dim, interleave=interleave, rope_scale=rope_scale,
                                             rope_theta=rope_theta, low_freq_factor=low_freq_factor,
                                             high_freq_factor=high_freq_factor, old_context_len=old_context_len)


@register_fake_op(f"{NAMESPACE}::apply_llama31_rope_pos_ids_inplace")
def _(q, k, pos_ids, rotary_dim=None, interleave=False, rope_scale=8.0, rope_theta=5e5, low_freq_factor=1.0, high_freq_factor=4.0,
      old_context_len=8192):
    return None


# ------------------------------------------------------------------ attention state merge, sampling
@register_custom_op(f"{NAMESPACE}::merge_state", mutates_args=())
def merge_state(v_a: torch.Tensor, s_a: torch.Tensor, v_b: torch.Tensor, s_b: torch.Tensor) -> List[torch.Tensor]:
    v, s = _cascade.merge_state(v_a, s_a, v_b, s_b)
    return [v, s]


@register_fake_op(f"{NAMESPACE}::merge_state")
def _(v_a, s_a, v_b, s_b):
    return [torch.empty_like(v_a), torch.empty_like(s_a)]


@register_custom_op(f"{NAMESPACE}::softmax", mutates_args=())
def softmax(logits: torch.Tensor, temperature: float = 1.0) -> torch.Tensor:
    return _sampling.softmax(logits, temperature)


@register_fake_op(f"{NAMESPACE}::softmax")
def _(logits, temperature=1.0):
    return torch.empty_like(logits, dtype=torch.float32)


@register_custom_op(f"{NAMESPACE}::top_k_renorm_probs", mutates_args=())
def top_k_renorm_probs(probs: torch.Tensor, top_k: int) -> torch.Tensor:
    return _sampling.top_k_renorm_probs(probs, top_k)


@register_fake_op(f"{NAMESPACE}::top_k_renorm_probs")
def _(probs, top_k):
    return torch.empty_like(probs)


@register_custom_op(f"{NAMESPACE}::sampling_from_probs", mutates_args=())
def sampling_from_probs(probs: torch.Tensor, seed: int, offset: int, deterministic: bool = True) -> torch.Tensor:
    return _sampling.sampling_from_probs(probs, deterministic=deterministic, seed=seed, offset=offset)


@register_fake_op(f"{NAMESPACE}::sampling_from_probs")
def _(probs, seed, offset, deterministic=True):
    return probs.new_empty(probs.shape[0], dtype=torch.int32)


REGISTERED = ("rmsnorm", "gemma_rmsnorm", "fused_add_rmsnorm", "silu_and_mul", "gelu_and_mul", "gelu_tanh_and_mul", "linear",
              "linear_gated_silu", "apply_rope_pos_ids_inplace", "apply_llama31_rope_pos_ids_inplace", "merge_state", "softmax",
              "top_k_renorm_probs", "sampling_from_probs")
