"""RMSNorm / LayerNorm family.  Parity: reference flashinfer/norm/__init__.py:112-640."""
from __future__ import annotations

from typing import Optional, Union

import torch

from . import jit, reference
from .utils import device_support_pdl, dtype_code, stream_ptr


def _pdl(enable_pdl, t):
    return 1 if (device_support_pdl(t.device) if enable_pdl is None else enable_pdl) else 0


def _strides3(t: torch.Tensor):
    """(rows, heads, s0, s1) for a 2-D [n, d] or 3-D [n, h, d] tensor with contiguous last dim."""
    if t.stride(-1) != 1:
        raise ValueError("last dim must be contiguous")
    if t.ndim == 2:
        return t.shape[0], 1, t.stride(0), 0
    if t.ndim == 3:
        return t.shape[0] * t.shape[1], t.shape[1], t.stride(0), t.stride(1)
    raise ValueError("expected a 2-D or 3-D tensor")


def _rms_launch(x, out, residual, weight, eps, weight_bias, scale, silu, enable_pdl):
    rows, heads, xs0, xs1 = _strides3(x)
    _, _, os0, os1 = _strides3(out)
    rs0 = rs1 = 0
    if residual is not None:
        _, _, rs0, rs1 = _strides3(residual)
    scale_ptr = scale if isinstance(scale, torch.Tensor) else None
    scale_val = 0.0 if isinstance(scale, torch.Tensor) or scale is None else float(scale)
    if scale_ptr is not None and scale_ptr.dtype != torch.float32:
        scale_ptr = scale_ptr.float()
    jit.load("norm").call(
        "rmsnorm_run", x, out, residual, weight, scale_ptr, rows, x.shape[-1], heads, xs0, xs1, os0, os1, rs0, rs1,
        float(eps), float(weight_bias), scale_val, 1 if silu else 0, dtype_code(x.dtype), dtype_code(out.dtype),
        _pdl(enable_pdl, x), stream_ptr(x),
    )


def rmsnorm(input: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6, out: Optional[torch.Tensor] = None,
            enable_pdl: Optional[bool] = None) -> torch.Tensor:
    """``out[i] = input[i] / RMS(input) * weight[i]`` for 2-D (tokens, hidden) or 3-D (tokens, heads, dim) input."""
    if out is None:
        out = torch.empty_like(input)
    if not input.is_cuda:
        out.copy_(reference.rmsnorm_ref(input, weight, eps))
        return out
    _rms_launch(input, out, None, weight, eps, 0.0, None, False, enable_pdl)
    return out


def gemma_rmsnorm(input, weight, eps: float = 1e-6, out=None, enable_pdl=None):
    """Gemma flavour: ``(1 + weight)`` scaling."""
    if out is None:
        out = torch.empty_like(input)
    if not input.is_cuda:
        out.copy_(reference.rmsnorm_ref(input, weight, eps, 1.0))
        return out
    _rms_launch(input, out, None, weight, eps, 1.0, None, False, enable_pdl)
    return out


def fused_add_rmsnorm(input: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6,
                      enable_pdl: Optional[bool] = None) -> None:
    """In place: ``residual += input; input = rmsnorm(residual) * weight``."""
    if not input.is_cuda:
        y, r = reference.fused_add_rmsnorm_ref(input, residual, weight, eps)
        residual.copy_(r)
        input.copy_(y)
        return
    _rms_launch(input, input, residual, weight, eps, 0.0, None, False, enable_pdl)


def gemma_fused_add_rmsnorm(input, residual, weight, eps: float = 1e-6, enable_pdl=None) -> None:
    if not input.is_cuda:
        y, r = reference.fused_add_rmsnorm_ref(input, residual, weight, eps, 1.0)
        residual.copy_(r)
        input.copy_(y)
        return
    _rms_launch(input, input, residual, weight, eps, 1.0, None, False, enable_pdl)


def _quant_ref(y, scale, dtype):
    s = scale.float().item() if isinstance(scale, torch.Tensor) else float(scale)
    lim = 448.0 if dtype == torch.float8_e4m3fn else 57344.0
    return (y.float() / s).clamp(-lim, lim).to(dtype)


def rmsnorm_quant(out: torch.Tensor, input: torch.Tensor, weight: torch.Tensor, scale: Union[float, torch.Tensor],
                  eps: float = 1e-6, enable_pdl: Optional[bool] = None) -> None:
    """``out = (rmsnorm(input) * weight / scale).to(fp8)``."""
    if not input.is_cuda:
        out.copy_(_quant_ref(reference.rmsnorm_ref(input.float(), weight, eps), scale, out.dtype))
        return
    _rms_launch(input, out, None, weight, eps, 0.0, scale, False, enable_pdl)


def fused_add_rmsnorm_quant(out, input, residual, weight, scale, eps: float = 1e-6, enable_pdl=None) -> None:
    """``residual += input; out = (rmsnorm(residual) * weight / scale).to(fp8)``."""
    if not input.is_cuda:
        r = (input.float() + residual.float()).to(input.dtype)
        residual.copy_(r)
        out.copy_(_quant_ref(reference.rmsnorm_ref(r.float(), weight, eps), scale, out.dtype))
        return
    _rms_launch(input, out, residual, weight, eps, 0.0, scale, False, enable_pdl)


def layernorm(input: torch.Tensor, gemma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """LayerNorm with fp32 gamma/beta (argument name ``gemma`` kept for reference parity)."""
    out = torch.empty_like(input)
    if not input.is_cuda:
        out.copy_(reference.layernorm_ref(input, gemma, beta, eps))
        return out
    jit.load("norm").call(
        "layernorm_run", input, out, gemma.float(), beta.float(), input.shape[0], input.shape[-1], input.stride(0),
        out.stride(0), float(eps), dtype_code(input.dtype), 1, stream_ptr(input),
    )
    return out


def fused_rmsnorm_silu(input, weight, eps: float = 1e-6, out=None, block_scale=None):
    """``out = SiLU(RMSNorm(input) * weight)`` (bf16 or fp8 output)."""
    if out is None:
        out = torch.empty_like(input)
    fp4_dt = getattr(torch, "float4_e2m1fn_x2", None)
    if out.dtype == torch.uint8 or (fp4_dt is not None and out.dtype == fp4_dt):
        # NVFP4 output (reference norm/__init__.py:639-658): linear [tokens, hidden / 16] e4m3 block scales, global scale 1.
        # Two native kernels: norm + SiLU, then the block quantiser.
        from .quantization.fp4 import fp4_quantize

        y = fused_rmsnorm_silu(input, weight, eps)
        q, sf = fp4_quantize(y, None, 16, False, False)
        out.view(torch.uint8).copy_(q)
        if block_scale is None:
            block_scale = sf.view(torch.float8_e4m3fn)
        else:
            block_scale.view(torch.uint8).copy_(sf)
        return out, block_scale
    if out.dtype not in (input.dtype, torch.float8_e4m3fn):
        raise TypeError(f"fused_rmsnorm_silu: unsupported output dtype {out.dtype}")
    if not input.is_cuda:
        y = reference.rmsnorm_ref(input.float(), weight, eps)
        out.copy_(torch.nn.functional.silu(y).to(out.dtype))
        return out
    _rms_launch(input, out, None, weight, eps, 0.0, None, True, None)
    return out


# ------------------------------------------------------------------ norm + block quantisation (reference cute_dsl/rmsnorm_fp4quant.py,
# add_rmsnorm_fp4quant.py) and DiT layernorm family (reference norm/__init__.py:1057-1440).  These are PDL-chained
# sequences of the native norm and quantisation kernels (one extra round trip through L2 compared with a single kernel).
def _norm_fp4quant(input, residual, weight, y_fp4, block_scale, global_scale, eps, block_size, scale_format, is_sf_swizzled_layout,
                   output_both_sf_layouts, enable_pdl):
    """One kernel: (residual += input) -> RMSNorm -> FP4 (csrc/elementwise/quantization.cu ``rmsnorm_fp4quant_kernel``)."""
    from .quantization.fp4 import _swizzled_sf_size, fp4_quantize

    h = input.shape[-1]
    ue8m0 = (scale_format == "ue8m0") or (scale_format is None and block_size == 32)
    x2 = input.reshape(-1, h)
    rows, kc = x2.shape[0], h // block_size
    native = (x2.is_cuda and x2.dtype in (torch.float16, torch.bfloat16) and weight.dtype == x2.dtype and h % block_size == 0
              and h % 8 == 0 and h * 4 <= 200 * 1024 and (residual is None or (residual.dtype == x2.dtype and residual.is_contiguous())))
    if not native:
        if residual is not None:
            x2 = x2.clone()
            fused_add_rmsnorm(x2, residual.reshape(-1, h), weight, eps, enable_pdl=enable_pdl)
            y = x2
        else:
            y = rmsnorm(x2, weight, eps, enable_pdl=enable_pdl)
        q, sf = fp4_quantize(y, global_scale, sf_vec_size=block_size, sf_use_ue8m0=ue8m0, is_sf_swizzled_layout=is_sf_swizzled_layout)
        sf_other = None
        if output_both_sf_layouts:
            _, sf_other = fp4_quantize(y, global_scale, sf_vec_size=block_size, sf_use_ue8m0=ue8m0,
                                       is_sf_swizzled_layout=not is_sf_swizzled_layout)
    else:
        if x2.stride(-1) != 1:
            x2 = x2.contiguous()
        q = torch.empty(rows, h // 2, dtype=torch.uint8, device=x2.device)
        sw_size = _swizzled_sf_size(rows, kc)

        def alloc(swizzled):
            return (torch.zeros(sw_size, dtype=torch.uint8, device=x2.device) if swizzled
                    else torch.empty(rows * kc, dtype=torch.uint8, device=x2.device))

        sf = alloc(is_sf_swizzled_layout)
        sf_other = alloc(not is_sf_swizzled_layout) if output_both_sf_layouts else None
        gs = global_scale.float().reshape(1).contiguous() if global_scale is not None else None
        r2 = residual.view(-1, h) if residual is not None else None
        pdl = device_support_pdl(x2.device) if enable_pdl is None else enable_pdl
        jit.load("quantization").call(
            "rmsnorm_fp4quant", x2, r2, weight.contiguous(), q, sf, sf_other, gs, rows, h, x2.stride(0), h, float(eps), block_size,
            1 if ue8m0 else 0, 1 if is_sf_swizzled_layout else 0, 0, dtype_code(x2.dtype), 1 if pdl else 0, stream_ptr(x2))
        sf = sf.view(-1, (kc + 3) // 4 * 4) if is_sf_swizzled_layout else sf.view(rows, kc)
        if sf_other is not None:
            sf_other = sf_other.view(rows, kc) if is_sf_swizzled_layout else sf_other.view(-1, (kc + 3) // 4 * 4)
    q = q.view(*input.shape[:-1], h // 2)
    if not is_sf_swizzled_layout:
        sf = sf.view(*input.shape[:-1], kc)
    if y_fp4 is not None:
        y_fp4.view(torch.uint8).copy_(q.view(torch.uint8))
        q = y_fp4
    if block_scale is not None:
        block_scale.view(torch.uint8).reshape(-1)[: sf.numel()].copy_(sf.view(torch.uint8).reshape(-1))
        sf = block_scale
    if output_both_sf_layouts:
        return q, sf, sf_other
    return q, sf


def rmsnorm_fp4quant(input: torch.Tensor, weight: torch.Tensor, y_fp4: Optional[torch.Tensor] = None,
                     block_scale: Optional[torch.Tensor] = None, global_scale: Optional[torch.Tensor] = None, eps: float = 1e-6,
                     block_size: int = 16, scale_format: Optional[str] = None, is_sf_swizzled_layout: bool = False,
                     enable_pdl: Optional[bool] = None):
    """``y = rmsnorm(input) * weight`` quantised to FP4 (NVFP4: block 16 / UE4M3, MXFP4: block 32 / UE8M0) in ONE kernel: the
    normalised activations never reach HBM.  Returns ``(y_fp4 [.., hidden/2] uint8, block_scale)``."""
    return _norm_fp4quant(input, None, weight, y_fp4, block_scale, global_scale, eps, block_size, scale_format,
                          is_sf_swizzled_layout, False, enable_pdl)


def add_rmsnorm_fp4quant(input: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor, y_fp4: Optional[torch.Tensor] = None,
                         block_scale: Optional[torch.Tensor] = None, global_scale: Optional[torch.Tensor] = None, eps: float = 1e-6,
                         block_size: int = 16, scale_format: Optional[str] = None, is_sf_swizzled_layout: bool = False,
                         output_both_sf_layouts: bool = False, enable_pdl: Optional[bool] = None):
    """``residual += input`` (in place), then :func:`rmsnorm_fp4quant` of the sum - one kernel.  With
    ``output_both_sf_layouts`` a third tensor holds the scales in the other (linear / swizzled) layout."""
    return _norm_fp4quant(input, residual, weight, y_fp4, block_scale, global_scale, eps, block_size, scale_format,
                          is_sf_swizzled_layout, output_both_sf_layouts, enable_pdl)


# DiT gate / residual / LayerNorm / modulation fusions live in diffusion_ops (re-exported here: the reference defines them in norm)
from .diffusion_ops import (  # noqa: E402,F401
    fused_dit_gate_residual_layernorm_gamma_beta,
    fused_dit_gate_residual_layernorm_scale_shift,
    fused_dit_residual_layernorm_scale_shift,
)


# ------------------------------------------------------------------ CuTe-DSL entry points of the reference (flashinfer/norm/__init__.py): same kernels here
rmsnorm_cute = rmsnorm
fused_add_rmsnorm_cute = fused_add_rmsnorm
rmsnorm_quant_cute = rmsnorm_quant
fused_add_rmsnorm_quant_cute = fused_add_rmsnorm_quant
layernorm_cute = layernorm


def qk_rmsnorm_cute(q: torch.Tensor, k: torch.Tensor, q_weight: torch.Tensor, k_weight: torch.Tensor, eps: float = 1e-6,
                    enable_pdl: Optional[bool] = None):
    """Per-head RMSNorm of q and k (``[..., heads, head_dim]``, normalised over ``head_dim``), in place like the reference."""
    for t, w in ((q, q_weight), (k, k_weight)):
        flat = t.reshape(-1, t.shape[-1])
        res = rmsnorm(flat, w, eps, enable_pdl=enable_pdl)
        if res.data_ptr() != flat.data_ptr():
            t.copy_(res.view_as(t))
    return q, k


from . import jit as _jit_acc  # noqa: E402

get_norm_module = _jit_acc.module_accessor("norm")
