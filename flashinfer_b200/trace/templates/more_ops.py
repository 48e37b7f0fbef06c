"""Norm + FP4, low-precision / masked / tiny GEMMs, NVFP4 quantisers, fused RoPE + fp8 + paged append, GDN multi-token decode,
Mamba-2 SSD prefill, decode context-parallel all-to-all (reference flashinfer/trace/templates/{norm,gemm,quantize,rope,gdn,mamba,
comm}.py - the entries that were still unbound here)."""
import math

import torch

from ..template import Const, Scalar, Tensor, TraceTemplate, Var
from .gdn import _AXES as _GDN_AXES
from .quantize import _fp4_quantize_reference, _x

_E2M1 = [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0]


def _dequant_nvfp4_rows(packed, sf_bytes, global_scale):
    """packed [M, K/2] bytes (low nibble first) + e4m3 scale bytes [M, K/16] -> fp32 [M, K] (divided by the global scale)."""
    grid = torch.tensor(_E2M1, device=packed.device)
    codes = torch.stack([packed & 0xF, packed >> 4], -1).reshape(packed.shape[0], -1).long()
    val = grid[codes & 7] * torch.where(codes >= 8, -1.0, 1.0)
    scale = sf_bytes.view(torch.uint8).view(torch.float8_e4m3fn).float().reshape(packed.shape[0], -1)
    return (val.view(packed.shape[0], -1, 16) * (scale / global_scale)[..., None]).reshape(packed.shape[0], -1)


def _rms(x, w, eps):
    x = x.float()
    return x * torch.rsqrt((x * x).mean(-1, keepdim=True) + eps) * w.float()


# ------------------------------------------------------------------ RMSNorm -> NVFP4 in one kernel
def _rmsnorm_fp4quant_reference(input, weight, global_scale, eps):
    """y = rmsnorm(input) * weight kept in fp32, then NVFP4: e4m3 scale per 16 elements = global_scale * amax / 6, e2m1 codes."""
    return _fp4_quantize_reference(_rms(input, weight, eps), global_scale, 16, False)


def _add_rmsnorm_fp4quant_reference(input, residual, weight, global_scale, eps):
    """residual' = residual + input (stored back in the residual's dtype), then the same as rmsnorm_fp4quant on residual'."""
    s = (input.float() + residual.float()).to(input.dtype)
    q, sf = _fp4_quantize_reference(_rms(s, weight, eps), global_scale, 16, False)
    return q, sf, s


def _norm_fp4_compare(residual_add):
    def compare(got, expected, kwargs):
        q, sf = got[0], got[1]
        x = kwargs["input"].float() + (kwargs["residual"].float() if residual_add else 0.0)
        if residual_add:
            x = x.to(kwargs["input"].dtype).float()
            torch.testing.assert_close(got[2].float(), expected[2].float(), atol=1e-2, rtol=1e-2)     # residual updated in place
        y = _rms(x, kwargs["weight"], kwargs["eps"])
        m, k = y.shape
        gs = float(kwargs["global_scale"])
        sf = sf.view(torch.uint8).reshape(m, k // 16)
        assert (sf.int() - expected[1].int()).abs().max() <= 1, "block scale bytes differ from the format definition by more than one code"
        deq = _dequant_nvfp4_rows(q.view(torch.uint8).reshape(m, k // 2), sf, gs)
        amax = y.view(m, k // 16, 16).abs().amax(-1, keepdim=True).expand(-1, -1, 16).reshape(m, k)
        assert ((deq - y).abs() <= 0.25 * amax + 0.02 * y.abs() + 1e-6).all(), "round trip exceeds the widest e2m1 half step"

    return compare


def _norm_fp4_init(*, M=64, K=4096, device="cuda", seed=0):
    x = _x(M, K, device, seed)
    g = torch.Generator(device="cpu").manual_seed(seed + 1)
    w = (1.0 + 0.1 * torch.randn(K, generator=g)).to(torch.bfloat16).to(device)
    return {"input": x, "weight": w, "global_scale": torch.tensor([448.0 * 6.0 / 8.0], device=device), "eps": 1e-6, "block_size": 16,
            "is_sf_swizzled_layout": False}


def _add_norm_fp4_init(*, M=64, K=4096, device="cuda", seed=0):
    kw = _norm_fp4_init(M=M, K=K, device=device, seed=seed)
    kw["residual"] = _x(M, K, device, seed + 7)
    return kw


_NORM_FP4_OUT = [Tensor("y_fp4", ("M", "K_half"), dtype="uint8"), Tensor("block_scale", ("M", "num_k_scales"), dtype="uint8")]
_NORM_FP4_CONSTRAINTS = ("K_half == K / 2", "num_k_scales == K / 16", "one == 1")

rmsnorm_fp4quant_trace = TraceTemplate(
    op_type="rmsnorm", name_fmt="rmsnorm_fp4quant_k{K}", axes=[Var("M"), Const("K")],
    inputs=[Tensor("input", ("M", "K")), Tensor("weight", ("K",)), Tensor("global_scale", ("one",), "float32"), Scalar("eps")],
    outputs=_NORM_FP4_OUT, reference=_rmsnorm_fp4quant_reference, init=_norm_fp4_init, compare=_norm_fp4_compare(False),
    helpers=(_rms, _fp4_quantize_reference), tags=("norm", "quantize", "nvfp4", "fused"), constraints=_NORM_FP4_CONSTRAINTS,
    description="RMSNorm with the NVFP4 block quantiser fused behind it: the normalised activations never reach HBM", test_sizes={"K": 128})

add_rmsnorm_fp4quant_trace = TraceTemplate(
    op_type="rmsnorm", name_fmt="add_rmsnorm_fp4quant_k{K}", axes=[Var("M"), Const("K")],
    inputs=[Tensor("input", ("M", "K")), Tensor("residual", ("M", "K")), Tensor("weight", ("K",)), Tensor("global_scale", ("one",), "float32"),
            Scalar("eps")],
    outputs=_NORM_FP4_OUT + [Tensor("residual_out", ("M", "K"), dtype_from="input", param="residual")],
    reference=_add_rmsnorm_fp4quant_reference, init=_add_norm_fp4_init, compare=_norm_fp4_compare(True), helpers=(_rms, _fp4_quantize_reference),
    tags=("norm", "quantize", "nvfp4", "fused", "inplace"), constraints=_NORM_FP4_CONSTRAINTS,
    description="Residual add (in place) + RMSNorm + NVFP4 block quantisation in one kernel", test_sizes={"K": 128})


def _fused_rmsnorm_silu_reference(input, weight, eps):
    y = _rms(input, weight, eps)
    return (y * torch.sigmoid(y)).to(input.dtype)


def _rmsnorm_silu_init(*, M=64, K=4096, device="cuda", seed=0):
    kw = _norm_fp4_init(M=M, K=K, device=device, seed=seed)
    return {"input": kw["input"], "weight": kw["weight"], "eps": 1e-6}


fused_rmsnorm_silu_trace = TraceTemplate(
    op_type="rmsnorm", name_fmt="fused_rmsnorm_silu_k{K}", axes=[Var("M"), Const("K")],
    inputs=[Tensor("input", ("M", "K")), Tensor("weight", ("K",)), Scalar("eps")], outputs=[Tensor("out", ("M", "K"), dtype_from="input")],
    reference=_fused_rmsnorm_silu_reference, init=_rmsnorm_silu_init, helpers=(_rms,), tags=("norm", "activation", "fused"), tolerance="bf16_norm",
    description="SiLU(RMSNorm(input) * weight) in one pass (bf16, fp8 or NVFP4 output)", test_sizes={"K": 128})


# ------------------------------------------------------------------ GEMMs
def _bmm_mxfp8_reference(A, B, A_scale, B_scale):
    """A [b, m, k] e4m3 with UE8M0 scales [b, m, k/32]; B [b, k, n] e4m3 (column-major storage) with scales [b, k/32, n]."""
    a = A.float() * torch.exp2(A_scale.view(torch.uint8).float() - 127).repeat_interleave(32, -1)[..., : A.shape[-1]]
    b = B.float() * torch.exp2(B_scale.view(torch.uint8).float() - 127).repeat_interleave(32, -2)[..., : B.shape[-2], :]
    return torch.bmm(a, b).to(torch.bfloat16)


def _mx_quant(x):
    *lead, k = x.shape
    g = x.float().reshape(*lead, k // 32, 32)
    exp = torch.ceil(torch.log2(g.abs().amax(-1).clamp_min(2.0 ** -100) / 448.0)).clamp(-127, 127)
    q = (g / torch.exp2(exp)[..., None]).clamp(-448, 448).reshape(*lead, k).to(torch.float8_e4m3fn)
    return q, (exp + 127).to(torch.uint8)


def _bmm_mxfp8_init(*, batch=4, m=128, n=256, k=512, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    a, sa = _mx_quant(torch.randn(batch, m, k, generator=g))
    bt, sbt = _mx_quant(torch.randn(batch, n, k, generator=g))               # quantised along k, stored [b, n, k]
    return {"A": a.to(device), "B": bt.transpose(1, 2).to(device), "A_scale": sa.to(device), "B_scale": sbt.transpose(1, 2).to(device),
            "dtype": torch.bfloat16}


bmm_mxfp8_trace = TraceTemplate(
    op_type="gemm", name_fmt="bmm_mxfp8_n{n}_k{k}", axes=[Var("batch"), Var("m"), Const("n"), Const("k"), Var("k_groups")],
    inputs=[Tensor("A", ("batch", "m", "k"), "float8_e4m3fn"), Tensor("B", ("batch", "k", "n"), "float8_e4m3fn"),
            Tensor("A_scale", ("batch", "m", "k_groups"), "uint8"), Tensor("B_scale", ("batch", "k_groups", "n"), "uint8")],
    outputs=[Tensor("out", ("batch", "m", "n"), dtype="bfloat16")], reference=_bmm_mxfp8_reference, init=_bmm_mxfp8_init, tags=("gemm", "mxfp8", "batched"),
    constraints=("k_groups == k / 32",), tolerance="cos", description="Batched MXFP8 GEMM (e4m3 data, power-of-two scale per 32 elements along k)",
    test_sizes={"n": 64, "k": 128, "m": 16, "batch": 2})


def _dq_1x128(x, s):
    return x.float() * s.float().repeat_interleave(128, -1)[..., : x.shape[-1]]


def _dq_128x128(w, s):
    return w.float() * s.float().repeat_interleave(128, -2).repeat_interleave(128, -1)[..., : w.shape[-2], : w.shape[-1]]


def _batch_deepgemm_reference(a, b, a_scale, b_scale, masked_m):
    """DeepGEMM masked layout: a [G, M_max, K] e4m3 + 1 x 128 scales, b [G, N, K] e4m3 + 128 x 128 scales; out[g, :masked_m[g]] =
    a[g] @ b[g]^T, rows past masked_m[g] are unspecified (zero here)."""
    out = torch.zeros(a.shape[0], a.shape[1], b.shape[1], dtype=torch.float32, device=a.device)
    for g in range(a.shape[0]):
        n = int(masked_m[g])
        out[g, :n] = _dq_1x128(a[g, :n], a_scale[g, :n]) @ _dq_128x128(b[g], b_scale[g]).t()
    return out.to(torch.bfloat16)


def _quant_1x128(x):
    *lead, k = x.shape
    g = x.float().reshape(*lead, k // 128, 128)
    s = g.abs().amax(-1).clamp_min(1e-6) / 448.0
    return (g / s[..., None]).reshape(*lead, k).to(torch.float8_e4m3fn), s.float()


def _quant_128x128(w):
    e, n, k = w.shape
    blocks = w.float().view(e, n // 128, 128, k // 128, 128)
    s = blocks.abs().amax((2, 4)).clamp_min(1e-6) / 448.0
    return (blocks / s[:, :, None, :, None]).to(torch.float8_e4m3fn).view(e, n, k), s.float()


def _batch_deepgemm_init(*, num_groups=4, max_m=256, n=512, k=512, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    a, sa = _quant_1x128(torch.randn(num_groups, max_m, k, generator=g))
    b, sb = _quant_128x128(torch.randn(num_groups, n, k, generator=g))
    masked = torch.randint(1, max_m + 1, (num_groups,), generator=g).int()
    return {"a": a.to(device), "b": b.to(device), "a_scale": sa.to(device), "b_scale": sb.to(device), "masked_m": masked.to(device), "expected_m": int(masked.float().mean())}


def _masked_rows_compare(got, expected, kwargs):
    out, ref, masked = got[0], expected[0], kwargs["masked_m"]
    for g in range(out.shape[0]):
        n = int(masked[g])
        o, r = out[g, :n].float(), ref[g, :n].float()
        assert torch.nn.functional.cosine_similarity(o.flatten(), r.flatten(), dim=0) > 0.99
        torch.testing.assert_close(o, r, atol=0.05 * float(r.abs().max()) + 1e-3, rtol=0.05)


batch_deepgemm_fp8_nt_groupwise_trace = TraceTemplate(
    op_type="gemm", name_fmt="batch_deepgemm_fp8_nt_groupwise_n{n}_k{k}", axes=[Var("num_groups"), Var("max_m"), Const("n"), Const("k"), Var("k_blocks"), Var("n_blocks")],
    inputs=[Tensor("a", ("num_groups", "max_m", "k"), "float8_e4m3fn"), Tensor("b", ("num_groups", "n", "k"), "float8_e4m3fn"),
            Tensor("a_scale", ("num_groups", "max_m", "k_blocks"), "float32"), Tensor("b_scale", ("num_groups", "n_blocks", "k_blocks"), "float32"),
            Tensor("masked_m", ("num_groups",), "int32")],
    outputs=[Tensor("out", ("num_groups", "max_m", "n"), dtype="bfloat16")], reference=_batch_deepgemm_reference, init=_batch_deepgemm_init,
    compare=_masked_rows_compare, helpers=(_dq_1x128, _dq_128x128), tags=("gemm", "fp8", "grouped", "masked"),
    constraints=("k_blocks == k / 128", "n_blocks == n / 128"),
    description="DeepGEMM masked grouped fp8 GEMM (1 x 128 activation scales, 128 x 128 weight scales); only the first masked_m[g] rows of a group count",
    test_sizes={"n": 128, "k": 256, "max_m": 128, "num_groups": 3})


def _grouped_gemm_nt_masked_reference(a, b, masked_m):
    out = torch.zeros(a.shape[0], a.shape[1], b.shape[1], dtype=torch.float32, device=a.device)
    for g in range(a.shape[0]):
        n = int(masked_m[g])
        out[g, :n] = a[g, :n].float() @ b[g].float().t()
    return out.to(a.dtype)


def _grouped_gemm_nt_masked_init(*, num_groups=4, max_m=256, n=512, k=512, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = (torch.randn(num_groups, max_m, k, generator=g) / k ** 0.25).to(torch.bfloat16)
    b = (torch.randn(num_groups, n, k, generator=g) / k ** 0.25).to(torch.bfloat16)
    return {"a": a.to(device), "b": b.to(device), "out": torch.zeros(num_groups, max_m, n, dtype=torch.bfloat16, device=device),
            "masked_m": torch.randint(1, max_m + 1, (num_groups,), generator=g).int().to(device)}


grouped_gemm_nt_masked_trace = TraceTemplate(
    op_type="gemm", name_fmt="grouped_gemm_nt_masked_n{n}_k{k}", axes=[Var("num_groups"), Var("max_m"), Const("n"), Const("k")],
    inputs=[Tensor("a", ("num_groups", "max_m", "k")), Tensor("b", ("num_groups", "n", "k")), Tensor("masked_m", ("num_groups",), "int32")],
    outputs=[Tensor("out", ("num_groups", "max_m", "n"), dtype_from="a", param="out")], reference=_grouped_gemm_nt_masked_reference,
    init=_grouped_gemm_nt_masked_init, compare=_masked_rows_compare, tags=("gemm", "bf16", "grouped", "masked"),
    description="Masked grouped GEMM out[g, :masked_m[g]] = a[g] @ b[g]^T (bf16 / fp16)", test_sizes={"n": 64, "k": 64, "max_m": 128, "num_groups": 3})


def _tinygemm_reference(input, weight, bias=None):
    y = input.float() @ weight.float().t()
    return (y + bias.float() if bias is not None else y).to(input.dtype)


def _tinygemm_init(*, m=4, n=256, k=7168, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return {"input": (torch.randn(m, k, generator=g) / k ** 0.25).to(torch.bfloat16).to(device),
            "weight": (torch.randn(n, k, generator=g) / k ** 0.25).to(torch.bfloat16).to(device),
            "bias": torch.randn(n, generator=g).to(torch.bfloat16).to(device)}


tinygemm_bf16_trace = TraceTemplate(
    op_type="gemm", name_fmt="tinygemm_bf16_n{n}_k{k}", axes=[Var("m"), Const("n"), Const("k")],
    inputs=[Tensor("input", ("m", "k")), Tensor("weight", ("n", "k")), Tensor("bias", ("n",), optional=True)],
    outputs=[Tensor("out", ("m", "n"), dtype_from="input")], reference=_tinygemm_reference, init=_tinygemm_init, tags=("gemm", "bf16", "small-m"),
    tolerance="cos", description="out = input @ weight^T + bias for a handful of rows (router / gate projections)", test_sizes={"n": 64, "k": 256, "m": 3})


# ------------------------------------------------------------------ NVFP4 quantisers
def _nvfp4_quantize_reference(a, a_global_sf):
    return _fp4_quantize_reference(a, a_global_sf, 16, False)


def _nvfp4_quantize_init(*, M=64, K=4096, device="cuda", seed=0):
    from ...quantization.fp4 import SfLayout

    x = _x(M, K, device, seed)
    return {"a": x, "a_global_sf": ((448.0 * 6.0) / x.float().abs().max()).reshape(1).to(device), "sfLayout": SfLayout.layout_linear}


def _nvfp4_quantize_compare(got, expected, kwargs):
    from .quantize import _fp4_compare_factory

    _fp4_compare_factory(16, False)(got, expected, {"input": kwargs["a"], "global_scale": kwargs["a_global_sf"]})


nvfp4_quantize_trace = TraceTemplate(
    op_type="quantize", name_fmt="nvfp4_quantize_rows_k{K}", axes=[Var("M"), Const("K")],
    inputs=[Tensor("a", ("M", "K")), Tensor("a_global_sf", ("one",), "float32")],
    outputs=[Tensor("packed", ("M", "K_half"), dtype="uint8"), Tensor("scale", ("M", "num_k_scales"), dtype="uint8")],
    reference=_nvfp4_quantize_reference, init=_nvfp4_quantize_init, compare=_nvfp4_quantize_compare, helpers=(_fp4_quantize_reference,),
    tags=("quantize", "nvfp4"), constraints=("K_half == K / 2", "num_k_scales == K / 16", "one == 1"),
    description="NVFP4 quantisation entry point with a selectable scale-factor layout (linear here; 128x4 / 8x4 swizzles re-order the same bytes)",
    test_sizes={"K": 128})


def _nvfp4_kv_reference(input, global_scale):
    """KV-cache flavour: input [M, K] (rows = tokens x heads); same block format, scales kept in the linear layout."""
    return _fp4_quantize_reference(input, global_scale, 16, False)


def _nvfp4_kv_init(*, M=64, K=128, device="cuda", seed=0):
    x = _x(M, K, device, seed)
    return {"input": x, "global_scale": ((448.0 * 6.0) / x.float().abs().max()).reshape(1).to(device)}


def _nvfp4_kv_compare(got, expected, kwargs):
    from .quantize import _fp4_compare_factory

    _fp4_compare_factory(16, False)(got, expected, kwargs)


nvfp4_kv_quantize_trace = TraceTemplate(
    op_type="quantize", name_fmt="nvfp4_kv_quantize_k{K}", axes=[Var("M"), Const("K")],
    inputs=[Tensor("input", ("M", "K")), Tensor("global_scale", ("one",), "float32")],
    outputs=[Tensor("packed", ("M", "K_half"), dtype="uint8"), Tensor("scale", ("M", "num_k_scales"), dtype="uint8")],
    reference=_nvfp4_kv_reference, init=_nvfp4_kv_init, compare=_nvfp4_kv_compare, helpers=(_fp4_quantize_reference,), tags=("quantize", "nvfp4", "kv-cache"),
    constraints=("K_half == K / 2", "num_k_scales == K / 16", "one == 1"),
    description="NVFP4 quantisation of K / V rows for an NVFP4 KV cache (linear scale layout)", test_sizes={"K": 128})


def _silu_mul_experts_quantize_reference(a, mask, a_global_sf):
    """a [E, M, 2K]: act = silu(a[..., :K]) * a[..., K:] (rounded to the input dtype), NVFP4 per expert with global scale a_global_sf[e];
    rows >= mask[e] are padding.  Scales are returned in the LINEAR layout [E, M, K / 16] (the op emits the 128x4 swizzle of the same bytes)."""
    e, m, k2 = a.shape
    k = k2 // 2
    act = (torch.nn.functional.silu(a[..., :k].float()) * a[..., k:].float()).to(a.dtype)
    outs = [_fp4_quantize_reference(act[i], a_global_sf.reshape(-1)[i], 16, False) for i in range(e)]
    return torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs])


def _silu_mul_experts_quantize_init(*, num_experts=4, M=128, K=256, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = torch.randn(num_experts, M, 2 * K, generator=g).to(torch.bfloat16)
    act = torch.nn.functional.silu(a[..., :K].float()) * a[..., K:].float()
    gs = (448.0 * 6.0) / act.abs().amax((1, 2))
    return {"a": a.to(device), "mask": torch.randint(1, M + 1, (num_experts,), generator=g).int().to(device), "a_global_sf": gs.float().to(device)}


def _silu_mul_experts_quantize_compare(got, expected, kwargs):
    from ...quantization.fp4 import e2m1_and_ufp8sf_scale_to_float

    q, sf = got[0], got[1]
    a, mask, gs = kwargs["a"], kwargs["mask"], kwargs["a_global_sf"].reshape(-1)
    e, m, k2 = a.shape
    k = k2 // 2
    act = (torch.nn.functional.silu(a[..., :k].float()) * a[..., k:].float()).to(a.dtype).float()
    assert tuple(q.shape) == (e, m, k // 2)
    for i in range(e):
        n = int(mask[i])
        deq = e2m1_and_ufp8sf_scale_to_float(q[i].view(torch.uint8), sf[i], gs[i].reshape(1), 16, 1, True)[:n].to(act.device)
        y = act[i, :n]
        amax = y.view(n, k // 16, 16).abs().amax(-1, keepdim=True).expand(-1, -1, 16).reshape(n, k)
        assert ((deq - y).abs() <= 0.25 * amax + 0.02 * y.abs() + 1e-6).all(), "round trip exceeds the widest e2m1 half step"


silu_and_mul_scaled_nvfp4_experts_quantize_trace = TraceTemplate(
    op_type="quantize", name_fmt="silu_and_mul_scaled_nvfp4_experts_quantize_k{K}", axes=[Var("num_experts"), Var("M"), Const("K"), Var("two_K")],
    inputs=[Tensor("a", ("num_experts", "M", "two_K")), Tensor("mask", ("num_experts",), "int32"), Tensor("a_global_sf", ("num_experts",), "float32")],
    outputs=[Tensor("packed", ("num_experts", "M", "K_half"), dtype="uint8"), Tensor("scale", ("num_experts", "swizzled_sf_bytes"), dtype="uint8")],
    reference=_silu_mul_experts_quantize_reference, init=_silu_mul_experts_quantize_init, compare=_silu_mul_experts_quantize_compare,
    helpers=(_fp4_quantize_reference,), derive=lambda sizes: {"K": sizes["two_K"] // 2} if "two_K" in sizes else {},
    tags=("quantize", "nvfp4", "moe", "activation", "fused"), constraints=("two_K == 2 * K", "K_half == K / 2", "swizzled_sf_bytes == round_up(M, 128) * round_up(K / 16, 4)"),
    description="Masked per-expert SwiGLU activation followed by NVFP4 quantisation (FC2 input of a masked-layout MoE), swizzled scale factors",
    test_sizes={"K": 64, "M": 128, "num_experts": 3})


# ------------------------------------------------------------------ GDN multi-token (speculative) decode
def _gdn_mtp_reference(q, k, v, initial_state, initial_state_indices, A_log, a, dt_bias, b, scale=None, use_qk_l2norm=True, disable_state_update=False):
    """q, k [B, T, H, K]; v [B, T, HV, V]; state pool [pool, HV, V, K] fp32 (K-last) addressed by initial_state_indices [B]; a, b [B, T, HV].
    Token t sees the state left by token t - 1; the pool row is updated with the state after the last token unless disable_state_update."""
    bsz, t_, h, kd = q.shape
    hv = v.shape[2]
    rep = hv // h
    sc = scale if scale is not None else kd ** -0.5
    g = torch.exp(-torch.exp(A_log.float()) * torch.nn.functional.softplus(a.float() + dt_bias.float()))
    beta = torch.sigmoid(b.float())
    out = torch.zeros(bsz, t_, hv, v.shape[3], dtype=torch.float32, device=q.device)
    pool = initial_state.float().clone()
    for i in range(bsz):
        s = pool[int(initial_state_indices[i])].transpose(-1, -2).clone()            # [HV, K, V]
        for t in range(t_):
            qt = q[i, t].float().repeat_interleave(rep, 0)
            kt = k[i, t].float().repeat_interleave(rep, 0)
            if use_qk_l2norm:
                qt = qt * torch.rsqrt((qt * qt).sum(-1, keepdim=True) + 1e-6)
                kt = kt * torch.rsqrt((kt * kt).sum(-1, keepdim=True) + 1e-6)
            s = s * g[i, t][:, None, None]
            delta = (v[i, t].float() - torch.einsum("hk,hkv->hv", kt, s)) * beta[i, t][:, None]
            s = s + kt[:, :, None] * delta[:, None, :]
            out[i, t] = torch.einsum("hk,hkv->hv", qt * sc, s)
        if not disable_state_update:
            pool[int(initial_state_indices[i])] = s.transpose(-1, -2)
    return out.to(q.dtype), pool


def _gdn_mtp_init(*, batch_size=4, num_tokens=4, num_q_heads=16, num_v_heads=32, head_dim_k=128, head_dim_v=128, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    bf = lambda t: t.to(torch.bfloat16).to(device)  # noqa: E731
    pool = batch_size + 3
    return {"q": bf(r(batch_size, num_tokens, num_q_heads, head_dim_k)), "k": bf(r(batch_size, num_tokens, num_q_heads, head_dim_k)),
            "v": bf(r(batch_size, num_tokens, num_v_heads, head_dim_v)), "initial_state": (r(pool, num_v_heads, head_dim_v, head_dim_k) * 0.1).to(device),
            "initial_state_indices": torch.randperm(pool, generator=g)[:batch_size].int().to(device), "A_log": (r(num_v_heads) * 0.5).to(device),
            "a": bf(r(batch_size, num_tokens, num_v_heads)), "dt_bias": (r(num_v_heads) * 0.1).to(device), "b": bf(r(batch_size, num_tokens, num_v_heads)), "disable_state_update": False}


gdn_mtp_trace = TraceTemplate(
    op_type="gdn", name_fmt="gdn_mtp_h{num_q_heads}_hv{num_v_heads}_k{head_dim_k}_v{head_dim_v}", axes=[Var("batch_size"), Var("num_tokens"), Var("pool_size")] + list(_GDN_AXES),
    inputs=[Tensor("q", ("batch_size", "num_tokens", "num_q_heads", "head_dim_k")), Tensor("k", ("batch_size", "num_tokens", "num_q_heads", "head_dim_k")),
            Tensor("v", ("batch_size", "num_tokens", "num_v_heads", "head_dim_v")),
            Tensor("initial_state", ("pool_size", "num_v_heads", "head_dim_v", "head_dim_k"), "float32"), Tensor("initial_state_indices", ("batch_size",), "int32"),
            Tensor("A_log", ("num_v_heads",)), Tensor("a", ("batch_size", "num_tokens", "num_v_heads")), Tensor("dt_bias", ("num_v_heads",)),
            Tensor("b", ("batch_size", "num_tokens", "num_v_heads")), Scalar("scale", optional=True), Scalar("use_qk_l2norm", "bool", optional=True),
            Scalar("disable_state_update", "bool", optional=True)],
    outputs=[Tensor("output", ("batch_size", "num_tokens", "num_v_heads", "head_dim_v"), dtype_from="q"),
             Tensor("state_out", ("pool_size", "num_v_heads", "head_dim_v", "head_dim_k"), dtype="float32", param="initial_state")],
    reference=_gdn_mtp_reference, init=_gdn_mtp_init, tags=("gdn", "decode", "mtp", "inplace"), tolerance="bf16",
    description="Gated delta rule over several draft tokens per sequence (speculative verification), state pool updated in place",
    test_sizes={"num_q_heads": 2, "num_v_heads": 4, "head_dim_k": 16, "head_dim_v": 8})


# ------------------------------------------------------------------ Mamba-2 SSD chunked prefill
def _ssd_reference(x, dt, A, B, C, D, dt_bias, dt_softplus, chunk_size):
    """x [b, L, H, P]; dt [b, L, H]; A [H] fp32 (negative); B, C [b, L, G, N]; D [H].  Token recurrence
    s <- exp(dt A) s + dt x (x) B,  y = s . C + D x;  y is returned in the chunked layout [b, H, P, L / chunk, chunk]."""
    b, L, H, P = x.shape
    G = B.shape[2]
    rep = H // G
    s = torch.zeros(b, H, P, B.shape[3], dtype=torch.float32, device=x.device)
    ys = []
    for t in range(L):
        d = dt[:, t].float() + dt_bias.float()
        d = torch.nn.functional.softplus(d) if dt_softplus else d
        s = s * torch.exp(d * A.float())[..., None, None] + (d[..., None] * x[:, t].float())[..., None] * B[:, t].float().repeat_interleave(rep, 1)[:, :, None, :]
        ys.append((s * C[:, t].float().repeat_interleave(rep, 1)[:, :, None, :]).sum(-1) + x[:, t].float() * D.float()[:, None])
    y = torch.stack(ys, 1)                                                      # [b, L, H, P]
    return y.view(b, L // chunk_size, chunk_size, H, P).permute(0, 3, 4, 1, 2).to(x.dtype), s


def _ssd_init(*, batch=2, seq_len=512, nheads=8, headdim=64, dstate=128, ngroups=1, chunk_size=128, device="cuda", seed=0):
    from ...mamba.ssd_combined import SSDCombined

    g = torch.Generator(device="cpu").manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    bf = lambda t: t.to(torch.bfloat16).to(device)  # noqa: E731
    w = SSDCombined(chunk_size, nheads, headdim, dstate, ngroups, state_dtype=torch.float32)
    return {"self": w, "x": bf(r(batch, seq_len, nheads, headdim)), "dt": bf(r(batch, seq_len, nheads) * 0.5), "A": (-torch.rand(nheads, generator=g) - 0.5).to(device),
            "B": bf(r(batch, seq_len, ngroups, dstate) * 0.3), "C": bf(r(batch, seq_len, ngroups, dstate) * 0.3), "D": r(nheads).to(device),
            "dt_bias": (r(nheads) * 0.1).to(device), "dt_softplus": True}


def _ssd_compare(got, expected, kwargs):
    for g_, e_ in zip(got, expected):
        cos = torch.nn.functional.cosine_similarity(g_.float().flatten(), e_.float().flatten(), dim=0)
        assert g_.shape == e_.shape and cos > 0.995, f"cosine similarity {float(cos):.4f}"
        torch.testing.assert_close(g_.float(), e_.float(), atol=0.06 * float(e_.float().abs().max()), rtol=0.06)


selective_scan_ssd_prefill_trace = TraceTemplate(
    op_type="mamba", name_fmt="ssd_combined_h{nheads}_p{headdim}_n{dstate}_g{ngroups}_c{chunk_size}",
    axes=[Var("batch"), Var("seq_len"), Var("num_chunks"), Const("nheads", abbrev="h"), Const("headdim", abbrev="p"), Const("dstate", abbrev="n"),
          Const("ngroups", abbrev="g"), Const("chunk_size", abbrev="c")],
    inputs=[Tensor("x", ("batch", "seq_len", "nheads", "headdim")), Tensor("dt", ("batch", "seq_len", "nheads")), Tensor("A", ("nheads",), "float32"),
            Tensor("B", ("batch", "seq_len", "ngroups", "dstate")), Tensor("C", ("batch", "seq_len", "ngroups", "dstate")), Tensor("D", ("nheads",)),
            Tensor("dt_bias", ("nheads",)), Scalar("dt_softplus", "bool"), Scalar("chunk_size", "int32", param="self.chunk_size")],
    outputs=[Tensor("y", ("batch", "nheads", "headdim", "num_chunks", "chunk_size"), dtype_from="x"),
             Tensor("final_states", ("batch", "nheads", "headdim", "dstate"), dtype="float32")],
    reference=_ssd_reference, init=_ssd_init, compare=_ssd_compare, tags=("mamba", "ssd", "prefill", "chunked"),
    constraints=("num_chunks == seq_len / chunk_size",),
    description="Mamba-2 state-space duality prefill: chunk-local GEMMs + a recurrence over chunk states (reference mamba/ssd_combined.py)",
    test_sizes={"nheads": 4, "headdim": 8, "dstate": 16, "ngroups": 1, "chunk_size": 16, "seq_len": 64, "batch": 2})


# ------------------------------------------------------------------ fused RoPE + fp8 quantisation + paged KV append
def _rope_fp8_append_reference(q_rope, k_rope, q_nope, k_nope, v, cos_sin_cache, pos_ids, k_cache, v_cache, kv_indices, kv_indptr, batch_indices,
                               positions, quant_scale_q, quant_scale_kv):
    """GQA flavour: heads are stored [rope | nope]: q = [rope(q_rope) | q_nope], k = [rope(k_rope) | k_nope] (neox halves), scaled and cast to e4m3; k / v rows are
    written to page kv_indices[kv_indptr[b] + pos // page_size], slot pos % page_size of an NHD cache."""
    def rot(x, pos):
        half = x.shape[-1] // 2
        cs = cos_sin_cache[pos.long()].float()
        cos, sin = cs[:, None, :half], cs[:, None, half:]
        x1, x2 = x.float()[..., :half], x.float()[..., half:]
        return torch.cat([x1 * cos - x2 * sin, x2 * cos + x1 * sin], -1)

    f8 = torch.float8_e4m3fn
    q_r = (rot(q_rope, pos_ids) * quant_scale_q).to(f8)
    q_n = (q_nope.float() * quant_scale_q).to(f8)
    k_full = torch.cat([rot(k_rope, pos_ids), k_nope.float()], -1) * quant_scale_kv
    kc, vc = k_cache.clone(), v_cache.clone()
    page_size = kc.shape[1]
    for t in range(k_full.shape[0]):
        b, p = int(batch_indices[t]), int(positions[t])
        page = int(kv_indices[int(kv_indptr[b]) + p // page_size])
        kc[page, p % page_size] = k_full[t].to(f8)
        vc[page, p % page_size] = (v[t].float() * quant_scale_kv).to(f8)
    return q_r, q_n, kc, vc


def _rope_fp8_append_init(*, batch_size=4, num_qo_heads=32, num_kv_heads=8, rope_dim=64, nope_dim=64, page_size=16, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16).to(device)  # noqa: E731
    new = [int(x) for x in torch.randint(1, 6, (batch_size,), generator=g)]
    old = [int(x) for x in torch.randint(0, 3 * page_size, (batch_size,), generator=g)]
    per = [(o + n + page_size - 1) // page_size for o, n in zip(old, new)]
    total = sum(per)
    kv_indptr = torch.tensor([0] + list(torch.tensor(per).cumsum(0)), dtype=torch.int32)
    kv_indices = torch.randperm(total + 2, generator=g)[:total].int()
    batch_indices = torch.cat([torch.full((n,), i, dtype=torch.int32) for i, n in enumerate(new)])
    positions = torch.cat([torch.arange(o, o + n, dtype=torch.int32) for o, n in zip(old, new)])
    nnz = sum(new)
    max_pos = max(o + n for o, n in zip(old, new)) + 1
    inv = 1.0 / (10000.0 ** (torch.arange(0, rope_dim, 2).float() / rope_dim))
    ang = torch.arange(max_pos).float()[:, None] * inv[None, :]
    cache = torch.cat([ang.cos(), ang.sin()], -1)
    hd = rope_dim + nope_dim
    kc = torch.zeros(total + 2, page_size, num_kv_heads, hd, dtype=torch.float8_e4m3fn)
    return {"q_rope": r(nnz, num_qo_heads, rope_dim), "k_rope": r(nnz, num_kv_heads, rope_dim), "q_nope": r(nnz, num_qo_heads, nope_dim),
            "k_nope": r(nnz, num_kv_heads, nope_dim), "v": r(nnz, num_kv_heads, hd), "cos_sin_cache": cache.to(device), "pos_ids": positions.to(device),
            "paged_kv_cache": (kc.to(device), kc.clone().to(device)), "kv_indices": kv_indices.to(device), "kv_indptr": kv_indptr.to(device),
            "batch_indices": batch_indices.to(device), "positions": positions.to(device), "is_neox": True, "quantize_dtype": torch.float8_e4m3fn,
            "quant_scale_q": 0.5, "quant_scale_kv": 0.25, "page_size": page_size, "kv_layout": "NHD"}


def _rope_fp8_append_compare(got, expected, kwargs):
    for g_, e_ in zip(got, expected):
        torch.testing.assert_close(g_.float(), e_.float(), atol=0.07, rtol=0.13)          # one e4m3 step


_PAGED8 = ("num_pages", "page_size", "num_kv_heads", "head_dim")
rope_quantize_fp8_append_paged_kv_cache_trace = TraceTemplate(
    op_type="rope", name_fmt="rope_quantize_fp8_append_paged_kv_cache_h{num_qo_heads}_kv{num_kv_heads}_r{rope_dim}_n{nope_dim}_ps{page_size}",
    axes=[Var("nnz"), Var("max_pos"), Var("num_pages"), Var("len_indptr"), Var("num_kv_indices"), Var("head_dim"), Const("num_qo_heads", abbrev="h"),
          Const("num_kv_heads", abbrev="kv"), Const("rope_dim", abbrev="r"), Const("nope_dim", abbrev="n"), Const("page_size", abbrev="ps")],
    inputs=[Tensor("q_rope", ("nnz", "num_qo_heads", "rope_dim")), Tensor("k_rope", ("nnz", "num_kv_heads", "rope_dim")),
            Tensor("q_nope", ("nnz", "num_qo_heads", "nope_dim")), Tensor("k_nope", ("nnz", "num_kv_heads", "nope_dim")),
            Tensor("v", ("nnz", "num_kv_heads", "head_dim")), Tensor("cos_sin_cache", ("max_pos", "rope_dim"), "float32"), Tensor("pos_ids", ("nnz",), "int32"),
            Tensor("k_cache", _PAGED8, "float8_e4m3fn", param="paged_kv_cache", tuple_idx=0),
            Tensor("v_cache", _PAGED8, "float8_e4m3fn", param="paged_kv_cache", tuple_idx=1), Tensor("kv_indices", ("num_kv_indices",), "int32"),
            Tensor("kv_indptr", ("len_indptr",), "int32"), Tensor("batch_indices", ("nnz",), "int32"), Tensor("positions", ("nnz",), "int32"),
            Scalar("quant_scale_q"), Scalar("quant_scale_kv")],
    outputs=[Tensor("q_rope_out", ("nnz", "num_qo_heads", "rope_dim"), dtype="float8_e4m3fn"),
             Tensor("q_nope_out", ("nnz", "num_qo_heads", "nope_dim"), dtype="float8_e4m3fn"),
             Tensor("k_cache_out", _PAGED8, dtype="float8_e4m3fn", param="paged_kv_cache", tuple_idx=0),
             Tensor("v_cache_out", _PAGED8, dtype="float8_e4m3fn", param="paged_kv_cache", tuple_idx=1)],
    reference=_rope_fp8_append_reference, init=_rope_fp8_append_init, compare=_rope_fp8_append_compare, tags=("rope", "quantize", "fp8", "page", "fused"),
    constraints=("head_dim == rope_dim + nope_dim",),
    description="RoPE on the rotary slices of q / k, fp8 quantisation of q, k, v and the paged KV-cache append in one kernel",
    test_sizes={"num_qo_heads": 4, "num_kv_heads": 2, "rope_dim": 16, "nope_dim": 16, "page_size": 4})


# ------------------------------------------------------------------ comm (schema only: the op needs a process group)
decode_cp_a2a_alltoall_trace = TraceTemplate(
    op_type="comm", name_fmt="decode_cp_a2a_alltoall_cp{cp_size}_d{head_dim}",
    axes=[Var("batch_size"), Var("num_heads"), Const("cp_size", abbrev="cp"), Const("head_dim", abbrev="d"), Var("stats_dim")],
    inputs=[Tensor("partial_o", ("batch_size", "num_heads", "cp_size", "head_dim")), Tensor("softmax_stats", ("batch_size", "num_heads", "cp_size", "stats_dim"), "float32"),
            Scalar("cp_rank", "int32"), Scalar("cp_size", "int32")],
    outputs=[Tensor("recv_o", ("batch_size", "num_heads", "cp_size", "head_dim"), dtype_from="partial_o"),
             Tensor("recv_stats", ("batch_size", "num_heads", "cp_size", "stats_dim"), dtype="float32")],
    tags=("comm", "context-parallel", "all-to-all"),
    description="Decode context-parallel exchange: slice [..., j, :] of the partial outputs and softmax statistics goes to rank j (one kernel for both)")

__all__ = [n for n in dir() if n.endswith("_trace")]
