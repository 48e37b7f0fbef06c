"""Diffusion-transformer fused ops (reference flashinfer/diffusion_ops): gate / residual / LayerNorm / modulation fusions."""
from ..norm import (  # noqa: F401
    fused_dit_gate_residual_layernorm_gamma_beta,
    fused_dit_gate_residual_layernorm_scale_shift,
    fused_dit_residual_layernorm_scale_shift,
)

__all__ = [
    "fused_dit_gate_residual_layernorm_gamma_beta",
    "fused_dit_gate_residual_layernorm_scale_shift",
    "fused_dit_residual_layernorm_scale_shift",
]
