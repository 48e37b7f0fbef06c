"""Gated-activation templates (reference flashinfer/trace/templates/activation.py)."""
import torch

from ..template import Const, Tensor, TraceTemplate, Var

_AXES = [Var("num_tokens"), Const("hidden_size", abbrev="h", description="width of the OUTPUT; the input holds [gate | up] = 2x")]


def _init(*, num_tokens=8, hidden_size=4096, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return {"input": torch.randn(num_tokens, 2 * hidden_size, generator=g).to(torch.bfloat16).to(device)}


def _silu_and_mul_reference(input):
    d = input.shape[-1] // 2
    g, u = input[..., :d].to(torch.float32), input[..., d:].to(torch.float32)
    return (g * torch.sigmoid(g) * u).to(input.dtype)


def _gelu_and_mul_reference(input):
    d = input.shape[-1] // 2
    g, u = input[..., :d].to(torch.float32), input[..., d:].to(torch.float32)
    return (torch.nn.functional.gelu(g) * u).to(input.dtype)


def _gelu_tanh_and_mul_reference(input):
    d = input.shape[-1] // 2
    g, u = input[..., :d].to(torch.float32), input[..., d:].to(torch.float32)
    return (torch.nn.functional.gelu(g, approximate="tanh") * u).to(input.dtype)


def _make(name, ref, desc):
    return TraceTemplate(
        op_type="activation", name_fmt=name + "_h{hidden_size}", axes=_AXES,
        inputs=[Tensor("input", ("num_tokens", "gate_up_size"))],
        outputs=[Tensor("output", ("num_tokens", "hidden_size"), dtype_from="input")], reference=ref, init=_init,
        tags=("activation",), description=desc, constraints=("gate_up_size == 2 * hidden_size",), tolerance="bf16_norm",
        derive=lambda s: {"hidden_size": s["gate_up_size"] // 2} if "gate_up_size" in s else {})


silu_and_mul_trace = _make("silu_and_mul", _silu_and_mul_reference, "SwiGLU: silu(x[:, :d]) * x[:, d:]")
gelu_and_mul_trace = _make("gelu_and_mul", _gelu_and_mul_reference, "GeGLU with the exact (erf) GELU")
gelu_tanh_and_mul_trace = _make("gelu_tanh_and_mul", _gelu_tanh_and_mul_reference, "GeGLU with the tanh approximation")
