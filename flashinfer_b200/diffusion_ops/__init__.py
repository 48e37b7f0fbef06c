"""Diffusion-transformer (DiT) block fusions: gated residual update + LayerNorm + adaptive modulation, optionally followed by the NVFP4 /
MXFP8 block quantiser of the next GEMM (reference flashinfer/norm/__init__.py:1057-1440, re-exported by flashinfer/diffusion_ops).

    residual_out = residual + input * (gate + gate_bias)            (gated forms)
    norm_out     = LayerNorm(residual_out) * gamma + beta           (gamma / beta form)
                 = LayerNorm(residual_out) * (1 + scale) + shift    (adaLN scale / shift form)

The LayerNorm is the native kernel (csrc/elementwise/norm.cu); the gate / modulation arithmetic around it is elementwise and runs in
fp32 on the stream (a single fused kernel for the whole block is not written)."""
from __future__ import annotations

import torch


def _dit_finish(res: torch.Tensor, normed: torch.Tensor, use_nvfp4: bool, use_mxfp8: bool, global_scaling_factor, residual_out,
                norm_out, sf_out):
    if residual_out is not None:
        residual_out.copy_(res)
        res = residual_out
    if use_nvfp4:
        from ..quantization.fp4 import fp4_quantize

        q, sf = fp4_quantize(normed.reshape(-1, normed.shape[-1]), global_scaling_factor, 16, False, True)
        return res, q.view(*normed.shape[:-1], -1), sf
    if use_mxfp8:
        from ..quantization.fp8 import mxfp8_quantize

        q, sf = mxfp8_quantize(normed.reshape(-1, normed.shape[-1]), True)
        return res, q.view(normed.shape), sf
    if norm_out is not None:
        norm_out.copy_(normed)
        normed = norm_out
    return res, normed


def fused_dit_gate_residual_layernorm_gamma_beta(input, residual, gate, gamma, beta, *, gate_bias=None, epsilon: float = 1e-6,
                                                 use_nvfp4: bool = False, use_mxfp8: bool = False, global_scaling_factor=None,
                                                 input_global_scaling_factor=None, residual_out=None, norm_out=None, sf_out=None):
    """``residual_out = residual + input * (gate + gate_bias)``; ``norm_out = LayerNorm(residual_out, gamma, beta)``."""
    x = input.float() * (float(input_global_scaling_factor) if input_global_scaling_factor is not None else 1.0)
    g = gate.float() + (gate_bias.float() if gate_bias is not None else 0.0)
    res = (residual.float() + x * g).to(input.dtype)
    from ..norm import layernorm

    normed = layernorm(res.reshape(-1, res.shape[-1]), gamma.float(), beta.float(), epsilon).view(res.shape)
    return _dit_finish(res, normed, use_nvfp4, use_mxfp8, global_scaling_factor, residual_out, norm_out, sf_out)


def fused_dit_gate_residual_layernorm_scale_shift(input, residual, gate, scale, shift, *, gate_bias=None, scale_bias=None,
                                                  shift_bias=None, epsilon: float = 1e-6, use_nvfp4: bool = False,
                                                  use_mxfp8: bool = False, global_scaling_factor=None,
                                                  input_global_scaling_factor=None, residual_out=None, norm_out=None, sf_out=None):
    """``residual_out = residual + input * gate``; ``norm_out = LayerNorm(residual_out) * (1 + scale) + shift`` (adaLN)."""
    x = input.float() * (float(input_global_scaling_factor) if input_global_scaling_factor is not None else 1.0)
    g = gate.float() + (gate_bias.float() if gate_bias is not None else 0.0)
    res = (residual.float() + x * g).to(input.dtype)
    return _dit_scale_shift(res, scale, shift, scale_bias, shift_bias, epsilon, use_nvfp4, use_mxfp8, global_scaling_factor,
                            residual_out, norm_out, sf_out)


def fused_dit_residual_layernorm_scale_shift(input, residual, scale, shift, *, scale_bias=None, shift_bias=None,
                                             epsilon: float = 1e-6, use_nvfp4: bool = False, use_mxfp8: bool = False,
                                             global_scaling_factor=None, input_global_scaling_factor=None, residual_out=None,
                                             norm_out=None, sf_out=None):
    """``residual_out = residual + input``; ``norm_out = LayerNorm(residual_out) * (1 + scale) + shift``."""
    x = input.float() * (float(input_global_scaling_factor) if input_global_scaling_factor is not None else 1.0)
    res = (residual.float() + x).to(input.dtype)
    return _dit_scale_shift(res, scale, shift, scale_bias, shift_bias, epsilon, use_nvfp4, use_mxfp8, global_scaling_factor,
                            residual_out, norm_out, sf_out)


def _dit_scale_shift(res, scale, shift, scale_bias, shift_bias, epsilon, use_nvfp4, use_mxfp8, gsf, residual_out, norm_out, sf_out):
    from ..norm import layernorm

    h = res.shape[-1]
    ones = torch.ones(h, dtype=torch.float32, device=res.device)
    zeros = torch.zeros(h, dtype=torch.float32, device=res.device)
    ln = layernorm(res.reshape(-1, h), ones, zeros, epsilon).view(res.shape).float()
    sc = scale.float() + (scale_bias.float() if scale_bias is not None else 0.0)
    sh = shift.float() + (shift_bias.float() if shift_bias is not None else 0.0)
    normed = (ln * (1.0 + sc) + sh).to(res.dtype)
    return _dit_finish(res, normed, use_nvfp4, use_mxfp8, gsf, residual_out, norm_out, sf_out)


__all__ = ["fused_dit_gate_residual_layernorm_gamma_beta", "fused_dit_gate_residual_layernorm_scale_shift",
           "fused_dit_residual_layernorm_scale_shift"]
