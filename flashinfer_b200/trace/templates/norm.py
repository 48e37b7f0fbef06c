"""Normalisation templates (reference flashinfer/trace/templates/norm.py).  Every ``reference`` is self-contained
PyTorch: its source is embedded in the definition file and must run on its own."""
import torch

from ..template import Const, Scalar, Tensor, TraceTemplate, Var

_AXES = [Var("batch_size"), Const("hidden_size", abbrev="h")]
_X = ("batch_size", "hidden_size")


def _rows(batch_size, hidden_size, device, seed, dtype=torch.bfloat16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(batch_size, hidden_size, generator=g).to(dtype).to(device)
    w = (torch.randn(hidden_size, generator=g) * 0.5 + 1.0).to(dtype).to(device)
    return g, x, w


def _rmsnorm_reference(input, weight, eps=1e-6):
    x = input.to(torch.float32)
    y = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * weight.to(torch.float32)
    return y.to(input.dtype)


def _rmsnorm_init(*, batch_size=8, hidden_size=4096, device="cuda", seed=0):
    _, x, w = _rows(batch_size, hidden_size, device, seed)
    return {"input": x, "weight": w, "eps": 1e-6}


rmsnorm_trace = TraceTemplate(
    op_type="rmsnorm", name_fmt="rmsnorm_h{hidden_size}", axes=_AXES,
    inputs=[Tensor("input", _X), Tensor("weight", ("hidden_size",)), Scalar("eps", optional=True)],
    outputs=[Tensor("output", _X, dtype_from="input")], reference=_rmsnorm_reference, init=_rmsnorm_init, tags=("norm",),
    description="Root-mean-square normalisation: y = x / sqrt(mean(x^2) + eps) * weight", tolerance="bf16_norm")


def _fused_add_rmsnorm_reference(input, residual, weight, eps=1e-6):
    """In place in the API: residual <- input + residual (rounded to the storage dtype), input <- rmsnorm(residual)."""
    r = (input.to(torch.float32) + residual.to(torch.float32)).to(input.dtype)
    x = r.to(torch.float32)
    y = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * weight.to(torch.float32)
    return y.to(input.dtype), r


def _fused_add_rmsnorm_init(*, batch_size=8, hidden_size=5120, device="cuda", seed=0):
    g, x, w = _rows(batch_size, hidden_size, device, seed)
    r = torch.randn(batch_size, hidden_size, generator=g).to(torch.bfloat16).to(device)
    return {"input": x, "residual": r, "weight": w, "eps": 1e-6}


fused_add_rmsnorm_trace = TraceTemplate(
    op_type="rmsnorm", name_fmt="fused_add_rmsnorm_h{hidden_size}", axes=_AXES,
    inputs=[Tensor("input", _X), Tensor("residual", _X), Tensor("weight", ("hidden_size",)), Scalar("eps", optional=True)],
    outputs=[Tensor("output", _X, dtype_from="input", param="input", description="written over `input`"),
             Tensor("residual_out", _X, dtype_from="input", param="residual", description="written over `residual`")],
    reference=_fused_add_rmsnorm_reference, init=_fused_add_rmsnorm_init, tags=("norm", "inplace"),
    description="Residual add fused with RMSNorm, both results in place", tolerance="bf16_norm")


def _gemma_rmsnorm_reference(input, weight, eps=1e-6):
    x = input.to(torch.float32)
    y = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * (1.0 + weight.to(torch.float32))
    return y.to(input.dtype)


gemma_rmsnorm_trace = TraceTemplate(
    op_type="rmsnorm", name_fmt="gemma_rmsnorm_h{hidden_size}", axes=_AXES,
    inputs=[Tensor("input", _X), Tensor("weight", ("hidden_size",)), Scalar("eps", optional=True)],
    outputs=[Tensor("output", _X, dtype_from="input")], reference=_gemma_rmsnorm_reference, init=_rmsnorm_init,
    tags=("norm", "gemma"), description="Gemma RMSNorm: the learned scale is (1 + weight)", tolerance="bf16_norm")


def _gemma_fused_add_rmsnorm_reference(input, residual, weight, eps=1e-6):
    r = (input.to(torch.float32) + residual.to(torch.float32)).to(input.dtype)
    x = r.to(torch.float32)
    y = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * (1.0 + weight.to(torch.float32))
    return y.to(input.dtype), r


gemma_fused_add_rmsnorm_trace = TraceTemplate(
    op_type="rmsnorm", name_fmt="gemma_fused_add_rmsnorm_h{hidden_size}", axes=_AXES,
    inputs=[Tensor("input", _X), Tensor("residual", _X), Tensor("weight", ("hidden_size",)), Scalar("eps", optional=True)],
    outputs=[Tensor("output", _X, dtype_from="input", param="input"), Tensor("residual_out", _X, dtype_from="input", param="residual")],
    reference=_gemma_fused_add_rmsnorm_reference, init=_fused_add_rmsnorm_init, tags=("norm", "gemma", "inplace"),
    description="Gemma residual add + RMSNorm in place", tolerance="bf16_norm")


def _layernorm_reference(input, gemma, beta, eps=1e-6):
    x = input.to(torch.float32)
    mu = x.mean(-1, keepdim=True)
    var = (x - mu).pow(2).mean(-1, keepdim=True)
    return ((x - mu) * torch.rsqrt(var + eps) * gemma.to(torch.float32) + beta.to(torch.float32)).to(input.dtype)


def _layernorm_init(*, batch_size=8, hidden_size=4096, device="cuda", seed=0):
    g, x, _ = _rows(batch_size, hidden_size, device, seed)
    gamma = (torch.randn(hidden_size, generator=g) * 0.5 + 1.0).to(device)
    beta = (torch.randn(hidden_size, generator=g) * 0.1).to(device)
    return {"input": x, "gemma": gamma, "beta": beta, "eps": 1e-6}


layernorm_trace = TraceTemplate(
    op_type="layernorm", name_fmt="layernorm_h{hidden_size}", axes=_AXES,
    inputs=[Tensor("input", _X), Tensor("gemma", ("hidden_size",), description="scale (the API's spelling of gamma)"),
            Tensor("beta", ("hidden_size",)), Scalar("eps", optional=True)],
    outputs=[Tensor("output", _X, dtype_from="input")], reference=_layernorm_reference, init=_layernorm_init, tags=("norm",),
    description="LayerNorm with fp32 scale and shift over bf16 activations", tolerance="bf16_norm")


def _rmsnorm_quant_reference(input, weight, scale, eps=1e-6):
    x = input.to(torch.float32)
    y = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * weight.to(torch.float32)
    s = float(scale) if not isinstance(scale, torch.Tensor) else scale.to(torch.float32)
    return (y / s).clamp(-448.0, 448.0).to(torch.float8_e4m3fn)


def _rmsnorm_quant_init(*, batch_size=8, hidden_size=4096, device="cuda", seed=0):
    _, x, w = _rows(batch_size, hidden_size, device, seed)
    return {"out": torch.empty(batch_size, hidden_size, dtype=torch.float8_e4m3fn, device=device), "input": x, "weight": w,
            "scale": 0.05, "eps": 1e-6}


rmsnorm_quant_trace = TraceTemplate(
    op_type="rmsnorm", name_fmt="rmsnorm_quant_h{hidden_size}", axes=_AXES,
    inputs=[Tensor("input", _X), Tensor("weight", ("hidden_size",)), Scalar("scale"), Scalar("eps", optional=True)],
    outputs=[Tensor("out", _X, dtype="float8_e4m3fn", param="out")], reference=_rmsnorm_quant_reference, init=_rmsnorm_quant_init,
    tags=("norm", "fp8"), description="RMSNorm with a static per-tensor fp8 (e4m3) quantisation of the result: q = y / scale",
    tolerance="fp8_quant")


def _fused_add_rmsnorm_quant_reference(input, residual, weight, scale, eps=1e-6):
    r = (input.to(torch.float32) + residual.to(torch.float32)).to(input.dtype)
    x = r.to(torch.float32)
    y = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * weight.to(torch.float32)
    s = float(scale) if not isinstance(scale, torch.Tensor) else scale.to(torch.float32)
    return (y / s).clamp(-448.0, 448.0).to(torch.float8_e4m3fn), r


def _fused_add_rmsnorm_quant_init(*, batch_size=8, hidden_size=4096, device="cuda", seed=0):
    kw = _rmsnorm_quant_init(batch_size=batch_size, hidden_size=hidden_size, device=device, seed=seed)
    g = torch.Generator(device="cpu").manual_seed(seed + 1)
    kw["residual"] = torch.randn(batch_size, hidden_size, generator=g).to(torch.bfloat16).to(device)
    return kw


fused_add_rmsnorm_quant_trace = TraceTemplate(
    op_type="rmsnorm", name_fmt="fused_add_rmsnorm_quant_h{hidden_size}", axes=_AXES,
    inputs=[Tensor("input", _X), Tensor("residual", _X), Tensor("weight", ("hidden_size",)), Scalar("scale"),
            Scalar("eps", optional=True)],
    outputs=[Tensor("out", _X, dtype="float8_e4m3fn", param="out"), Tensor("residual_out", _X, dtype_from="input", param="residual")],
    reference=_fused_add_rmsnorm_quant_reference, init=_fused_add_rmsnorm_quant_init, tags=("norm", "fp8", "inplace"),
    description="Residual add + RMSNorm + static fp8 quantisation", tolerance="fp8_quant")
