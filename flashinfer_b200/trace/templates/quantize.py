"""Quantisation templates (reference flashinfer/trace/templates/quantize.py).  References implement the format
definitions; because ties and scale rounding may legitimately differ by one code between implementations, each template
carries a ``compare`` that checks (1) the scale bytes against the reference within one code and (2) the round trip
dequantise(API output) against the *original* input within the format's worst-case step."""
import torch

from ..template import Const, Scalar, Tensor, TraceTemplate, Var

_AXES = [Var("M"), Const("K")]


def _x(M, K, device, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(M, K, generator=g) * torch.exp(torch.randn(M, 1, generator=g))      # rows of different magnitude
    return x.to(torch.bfloat16).to(device)


# ---- MXFP8: e4m3 elements, one UE8M0 (power-of-two) scale per 32 elements
def _mxfp8_quantize_reference(input):
    m, k = input.shape
    x = input.to(torch.float32).view(m, k // 32, 32)
    amax = x.abs().amax(-1)
    exp = torch.ceil(torch.log2(torch.clamp(amax / 448.0, min=2.0 ** -127))).clamp(-127, 127)   # smallest 2^e with amax / 2^e <= 448
    q = (x / torch.exp2(exp)[..., None]).clamp(-448, 448).view(m, k).to(torch.float8_e4m3fn)
    return q, (exp + 127).to(torch.uint8).reshape(-1)


def _mxfp8_init(*, M=64, K=4096, device="cuda", seed=0):
    return {"input": _x(M, K, device, seed), "is_sf_swizzled_layout": False}


def _mxfp8_compare(got, expected, kwargs):
    (q, sf), (_, sf_ref) = got, expected
    x = kwargs["input"].to(torch.float32)
    m, k = x.shape
    assert q.dtype == torch.float8_e4m3fn and sf.dtype == torch.uint8 and sf.numel() == m * k // 32
    assert (sf.reshape(-1).int() - sf_ref.int()).abs().max() <= 1, "UE8M0 scale bytes differ from the format definition by more than one code"
    deq = (q.to(torch.float32).view(m, k // 32, 32) * torch.exp2(sf.view(m, k // 32).float() - 127)[..., None]).view(m, k)
    amax = x.view(m, k // 32, 32).abs().amax(-1, keepdim=True).expand(-1, -1, 32).reshape(m, k)
    assert ((deq - x).abs() <= 0.0625 * x.abs() + amax * 2.0 ** -9 + 1e-30).all(), "MXFP8 round trip exceeds half an e4m3 step"


mxfp8_quantize_trace = TraceTemplate(
    op_type="quantize", name_fmt="mxfp8_quantize_k{K}", axes=_AXES,
    inputs=[Tensor("input", ("M", "K")), Scalar("is_sf_swizzled_layout", "bool", optional=True)],
    outputs=[Tensor("q", ("M", "K"), dtype="float8_e4m3fn"), Tensor("scale", ("num_scales",), dtype="uint8")],
    reference=_mxfp8_quantize_reference, init=_mxfp8_init, compare=_mxfp8_compare, tags=("quantize", "mxfp8"),
    constraints=("num_scales == M * K / 32 (linear layout; the swizzled layout pads M to 128 and K/32 to 4)",),
    description="OCP MXFP8: e4m3 data with a power-of-two scale per 32 elements", test_sizes={"K": 128})


# ---- NVFP4 / MXFP4: e2m1 elements packed two per byte
def _fp4_quantize_reference(input, global_scale=None, sf_vec_size=16, sf_use_ue8m0=False):
    """NVFP4 (vec 16, e4m3 scale): sf = e4m3(global_scale * amax / 6); element = rn_e2m1(x * global_scale / sf).
    MXFP4 (vec 32, UE8M0 scale): sf = 2^ceil(log2(amax / 6))."""
    grid = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0], device=input.device)
    m, k = input.shape
    gs = float(global_scale) if global_scale is not None else 1.0
    x = input.to(torch.float32).view(m, k // sf_vec_size, sf_vec_size)
    amax = x.abs().amax(-1)
    if sf_use_ue8m0:
        exp = torch.ceil(torch.log2(torch.clamp(amax / 6.0 * gs, min=2.0 ** -127))).clamp(-127, 127)
        sf_val, sf_byte = torch.exp2(exp), (exp + 127).to(torch.uint8)
    else:
        s8 = (gs * amax / 6.0).to(torch.float8_e4m3fn)
        sf_val, sf_byte = s8.to(torch.float32), s8.view(torch.uint8)
    y = x * torch.where(sf_val > 0, gs / sf_val, torch.zeros_like(sf_val))[..., None]
    idx = (y.abs().clamp(max=6.0)[..., None] - grid).abs().argmin(-1)                    # nearest code (ties: lower index)
    code = (idx.to(torch.uint8) | ((y < 0).to(torch.uint8) << 3)).view(m, k)
    return code[:, 0::2] | (code[:, 1::2] << 4), sf_byte.reshape(m, k // sf_vec_size)


def _fp4_compare_factory(vec, ue8m0):
    def compare(got, expected, kwargs):
        (packed, sf), (_, sf_ref) = got, expected
        x = kwargs["input"].to(torch.float32)
        m, k = x.shape
        gs = float(kwargs["global_scale"]) if kwargs.get("global_scale") is not None else 1.0
        assert packed.dtype == torch.uint8 and tuple(packed.shape) == (m, k // 2)
        sf = sf.view(torch.uint8).reshape(m, k // vec)
        assert (sf.int() - sf_ref.int()).abs().max() <= 1, "block scale bytes differ from the format definition by more than one code"
        grid = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0])
        lo, hi = packed & 0xF, packed >> 4
        codes = torch.stack([lo, hi], dim=-1).reshape(m, k).long()
        val = grid[codes & 7] * torch.where(codes >= 8, -1.0, 1.0)
        scale = torch.exp2(sf.float() - 127) if ue8m0 else sf.view(torch.float8_e4m3fn).float()
        deq = (val.view(m, k // vec, vec) * (scale / gs)[..., None]).view(m, k)
        amax = x.view(m, k // vec, vec).abs().amax(-1, keepdim=True).expand(-1, -1, vec).reshape(m, k)
        bound = (0.5 if ue8m0 else 0.25) * amax + 1e-6     # half the widest e2m1 step (4 -> 6) at the block's scale (+ scale rounding)
        assert ((deq - x).abs() <= bound).all(), "FP4 round trip exceeds the widest e2m1 half step"

    return compare


def _nvfp4_init(*, M=64, K=4096, device="cuda", seed=0):
    x = _x(M, K, device, seed)
    gs = (448.0 * 6.0) / x.float().abs().max()
    return {"input": x, "global_scale": gs.reshape(1).to(device), "sf_vec_size": 16, "sf_use_ue8m0": False, "is_sf_swizzled_layout": False}


def _mxfp4_init(*, M=64, K=4096, device="cuda", seed=0):
    return {"input": _x(M, K, device, seed), "global_scale": None, "sf_vec_size": 32, "sf_use_ue8m0": True, "is_sf_swizzled_layout": False}


_FP4_IN = [Tensor("input", ("M", "K")), Tensor("global_scale", ("one",), "float32", optional=True), Scalar("sf_vec_size", "int32", optional=True),
           Scalar("sf_use_ue8m0", "bool", optional=True)]
_FP4_OUT = [Tensor("packed", ("M", "K_half"), dtype="uint8"), Tensor("scale", ("M", "num_k_scales"), dtype="uint8")]

fp4_quantize_trace = TraceTemplate(
    op_type="quantize", name_fmt="nvfp4_quantize_k{K}", axes=_AXES, inputs=_FP4_IN, outputs=_FP4_OUT,
    reference=_fp4_quantize_reference, init=_nvfp4_init, compare=_fp4_compare_factory(16, False), tags=("quantize", "nvfp4"),
    constraints=("K_half == K / 2", "num_k_scales == K / sf_vec_size", "one == 1"),
    description="NVFP4: e2m1 pairs + one e4m3 scale per 16 elements under a global fp32 scale", test_sizes={"K": 128})

mxfp4_quantize_trace = TraceTemplate(
    op_type="quantize", name_fmt="mxfp4_quantize_k{K}", axes=_AXES, inputs=_FP4_IN, outputs=_FP4_OUT,
    reference=_fp4_quantize_reference, init=_mxfp4_init, compare=_fp4_compare_factory(32, True), tags=("quantize", "mxfp4"),
    constraints=("K_half == K / 2", "num_k_scales == K / sf_vec_size", "one == 1"),
    description="OCP MXFP4: e2m1 pairs + one power-of-two scale per 32 elements", test_sizes={"K": 128})
