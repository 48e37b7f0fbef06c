"""GEMM templates (reference flashinfer/trace/templates/gemm.py).  Weights are passed column-major: ``b`` is the ``[K, N]``
view of an ``[N, K]`` row-major weight (``w.t()``), the layout the tcgen05 kernels read without a transpose."""
import torch

from ..template import Const, Scalar, Tensor, TraceTemplate, Var

_SIZES = {"N": 96, "K": 128}


def _ab(M, N, K, device, seed, batch=None):
    g = torch.Generator(device="cpu").manual_seed(seed)
    lead = () if batch is None else (batch,)
    a = (torch.randn(*lead, M, K, generator=g) * 0.5).to(torch.bfloat16).to(device)
    w = (torch.randn(*lead, N, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(device)
    return g, a, w


def _mm_bf16_reference(a, b, bias=None):
    out = a.to(torch.float32) @ b.to(torch.float32)
    if bias is not None:
        out = out + bias.to(torch.float32)
    return out.to(torch.bfloat16)


def _mm_bf16_init(*, M=64, N=4096, K=4096, device="cuda", seed=0):
    _, a, w = _ab(M, N, K, device, seed)
    return {"a": a, "b": w.t()}


mm_bf16_trace = TraceTemplate(
    op_type="gemm", name_fmt="mm_bf16_n{N}_k{K}", axes=[Var("M"), Const("N"), Const("K")],
    inputs=[Tensor("a", ("M", "K")), Tensor("b", ("K", "N"), description="column-major: w.t() of an [N, K] weight"),
            Tensor("bias", ("N",), optional=True)],
    outputs=[Tensor("out", ("M", "N"), dtype="bfloat16")], reference=_mm_bf16_reference, init=_mm_bf16_init, tags=("gemm", "bf16"),
    description="out = a @ b (+ bias), fp32 accumulation", tolerance="cos", test_sizes=_SIZES)


def _tgv_init(*, M=16, N=4096, K=4096, device="cuda", seed=0):
    g, a, w = _ab(M, N, K, device, seed)
    return {"a": a, "b": w.t(), "bias": torch.randn(N, generator=g).to(torch.bfloat16).to(device)}


tgv_gemm_sm100_trace = TraceTemplate(
    op_type="gemm", name_fmt="tgv_gemm_sm100_n{N}_k{K}", axes=[Var("M"), Const("N"), Const("K")],
    inputs=[Tensor("a", ("M", "K")), Tensor("b", ("K", "N")), Tensor("bias", ("N",), optional=True)],
    outputs=[Tensor("out", ("M", "N"), dtype="bfloat16")], reference=_mm_bf16_reference, init=_tgv_init, tags=("gemm", "bf16", "small_m"),
    description="Latency-oriented small-M GEMM with bias (decode projections)", tolerance="cos", test_sizes=_SIZES)


def _bmm_bf16_reference(A, B):
    return torch.bmm(A.to(torch.float32), B.to(torch.float32)).to(torch.bfloat16)


def _bmm_bf16_init(*, batch=4, M=64, N=1024, K=1024, device="cuda", seed=0):
    _, a, w = _ab(M, N, K, device, seed, batch)
    return {"A": a, "B": w.transpose(-1, -2)}


bmm_bf16_trace = TraceTemplate(
    op_type="gemm", name_fmt="bmm_bf16_n{N}_k{K}", axes=[Var("batch"), Var("M"), Const("N"), Const("K")],
    inputs=[Tensor("A", ("batch", "M", "K")), Tensor("B", ("batch", "K", "N"))], outputs=[Tensor("out", ("batch", "M", "N"), dtype="bfloat16")],
    reference=_bmm_bf16_reference, init=_bmm_bf16_init, tags=("gemm", "bf16", "batched"), description="Batched bf16 GEMM",
    tolerance="cos", test_sizes=_SIZES)


def _bmm_fp8_reference(A, B, A_scale, B_scale, dtype=torch.bfloat16):
    out = torch.bmm(A.to(torch.float32), B.to(torch.float32)) * (A_scale.to(torch.float32) * B_scale.to(torch.float32))
    return out.to(dtype)


def _bmm_fp8_init(*, batch=4, M=64, N=1024, K=1024, device="cuda", seed=0):
    _, a, w = _ab(M, N, K, device, seed, batch)
    sa, sb = a.float().abs().max() / 448.0, w.float().abs().max() / 448.0
    return {"A": (a.float() / sa).to(torch.float8_e4m3fn), "B": (w.float() / sb).to(torch.float8_e4m3fn).transpose(-1, -2),
            "A_scale": sa.reshape(1).to(device), "B_scale": sb.reshape(1).to(device), "dtype": torch.bfloat16}


bmm_fp8_trace = TraceTemplate(
    op_type="gemm", name_fmt="bmm_fp8_n{N}_k{K}", axes=[Var("batch"), Var("M"), Const("N"), Const("K")],
    inputs=[Tensor("A", ("batch", "M", "K")), Tensor("B", ("batch", "K", "N")), Tensor("A_scale", ("one",), "float32"),
            Tensor("B_scale", ("one",), "float32"), Scalar("dtype", "dtype")],
    outputs=[Tensor("out", ("batch", "M", "N"), dtype="bfloat16")], reference=_bmm_fp8_reference, init=_bmm_fp8_init,
    tags=("gemm", "fp8", "batched"), constraints=("one == 1",), description="Batched e4m3 GEMM with per-tensor de-quantisation scales",
    tolerance="cos", test_sizes=_SIZES)


def _mm_fp8_reference(a, b, alpha=None):
    out = a.to(torch.float32) @ b.to(torch.float32)
    if alpha is not None:
        out = out * alpha.to(torch.float32)
    return out.to(torch.bfloat16)


def _mm_fp8_init(*, M=64, N=4096, K=4096, device="cuda", seed=0):
    _, a, w = _ab(M, N, K, device, seed)
    sa, sb = a.float().abs().max() / 448.0, w.float().abs().max() / 448.0
    return {"a": (a.float() / sa).to(torch.float8_e4m3fn), "b": (w.float() / sb).to(torch.float8_e4m3fn).t(),
            "alpha": (sa * sb).reshape(1).to(device)}


mm_fp8_trace = TraceTemplate(
    op_type="gemm", name_fmt="mm_fp8_n{N}_k{K}", axes=[Var("M"), Const("N"), Const("K")],
    inputs=[Tensor("a", ("M", "K")), Tensor("b", ("K", "N")), Tensor("alpha", ("one",), "float32", optional=True)],
    outputs=[Tensor("out", ("M", "N"), dtype="bfloat16")], reference=_mm_fp8_reference, init=_mm_fp8_init, tags=("gemm", "fp8"),
    constraints=("one == 1",), description="e4m3 GEMM scaled by alpha", tolerance="cos", test_sizes=_SIZES)


def _gemm_fp8_nt_groupwise_reference(a, b, a_scale, b_scale):
    """a [M, K] e4m3 with scales a_scale [K/128, M] (1 x 128 groups, MN-major); b [N, K] e4m3 with b_scale [K/128, N/128]
    (128 x 128 blocks)."""
    m, k = a.shape
    n = b.shape[0]
    sa = a_scale.to(torch.float32).t().repeat_interleave(128, dim=1)[:, :k]
    sb = b_scale.to(torch.float32).t().repeat_interleave(128, dim=0).repeat_interleave(128, dim=1)[:n, :k]
    return ((a.to(torch.float32) * sa) @ (b.to(torch.float32) * sb).t()).to(torch.bfloat16)


def _groupwise_init(*, M=64, N=4096, K=4096, device="cuda", seed=0):
    _, a, w = _ab(M, N, K, device, seed)
    af = a.float().view(M, K // 128, 128)
    sa = af.abs().amax(-1).clamp(min=1e-4) / 448.0                               # [M, K/128]
    wf = w.float().view(N // 128, 128, K // 128, 128)
    sb = wf.abs().amax((1, 3)).clamp(min=1e-4) / 448.0                           # [N/128, K/128]
    return {"a": (af / sa[..., None]).view(M, K).to(torch.float8_e4m3fn), "b": (wf / sb[:, None, :, None]).view(N, K).to(torch.float8_e4m3fn),
            "a_scale": sa.t().contiguous().to(device), "b_scale": sb.t().contiguous().to(device), "scale_major_mode": "MN",
            "out_dtype": torch.bfloat16}


gemm_fp8_nt_groupwise_trace = TraceTemplate(
    op_type="gemm", name_fmt="gemm_fp8_nt_groupwise_n{N}_k{K}", axes=[Var("M"), Const("N"), Const("K")],
    inputs=[Tensor("a", ("M", "K")), Tensor("b", ("N", "K")), Tensor("a_scale", ("k_groups", "M"), "float32"),
            Tensor("b_scale", ("k_groups", "n_groups"), "float32")],
    outputs=[Tensor("out", ("M", "N"), dtype="bfloat16")], reference=_gemm_fp8_nt_groupwise_reference, init=_groupwise_init,
    tags=("gemm", "fp8", "groupwise"), constraints=("k_groups == K / 128", "n_groups == N / 128"),
    description="DeepSeek-style fp8 GEMM: 1x128 activation groups, 128x128 weight blocks, scales applied per K block",
    tolerance="cos", test_sizes={"N": 256, "K": 256})


def _segment_gemm_reference(x, weights, seg_indptr, weight_indices=None, weight_column_major=True):
    """Rows seg_indptr[i] : seg_indptr[i+1] use weight weight_indices[i] (default i); column-major weights are [n, d_out, d_in]."""
    d_out = weights.shape[1] if weight_column_major else weights.shape[2]
    out = torch.zeros(x.shape[0], d_out, dtype=torch.float32, device=x.device)
    for i in range(seg_indptr.numel() - 1):
        s, e = int(seg_indptr[i]), int(seg_indptr[i + 1])
        w = weights[int(weight_indices[i]) if weight_indices is not None else i].to(torch.float32)
        out[s:e] = x[s:e].to(torch.float32) @ (w.t() if weight_column_major else w)
    return out.to(x.dtype)


def _segment_gemm_init(*, batch_size=4, d_in=4096, d_out=4096, device="cuda", seed=0):
    import flashinfer_b200 as fi

    g = torch.Generator(device="cpu").manual_seed(seed)
    lens = torch.randint(1, 20, (batch_size,), generator=g)
    indptr = torch.cat([torch.zeros(1, dtype=torch.long), lens.cumsum(0)])
    x = (torch.randn(int(indptr[-1]), d_in, generator=g) * 0.5).to(torch.bfloat16).to(device)
    w = (torch.randn(batch_size, d_out, d_in, generator=g) / d_in ** 0.5).to(torch.bfloat16).to(device)
    return {"self": fi.SegmentGEMMWrapper(torch.empty(1 << 20, dtype=torch.uint8, device=device)), "x": x, "weights": w,
            "batch_size": batch_size, "weight_column_major": True, "seg_indptr": indptr.to(device)}


segment_gemm_trace = TraceTemplate(
    op_type="gemm", name_fmt="segment_gemm_in{d_in}_out{d_out}", axes=[Var("total_rows"), Var("batch_size"), Var("len_indptr"), Const("d_in"), Const("d_out")],
    inputs=[Tensor("x", ("total_rows", "d_in")), Tensor("weights", ("batch_size", "d_out", "d_in")),
            Tensor("seg_indptr", ("len_indptr",), optional=True), Tensor("weight_indices", ("batch_size",), optional=True),
            Scalar("weight_column_major", "bool")],
    outputs=[Tensor("y", ("total_rows", "d_out"), dtype_from="x")], reference=_segment_gemm_reference, init=_segment_gemm_init,
    tags=("gemm", "grouped"), constraints=("len_indptr == batch_size + 1",), description="Per-segment GEMM (LoRA / grouped experts)",
    tolerance="cos", test_sizes={"d_in": 64, "d_out": 48})


# ---- block-scaled FP4 GEMM (scales in the 128x4-tiled layout the tensor core's scale path reads)
def _mm_fp4_reference(a, b, a_descale, b_descale, alpha=None, block_size=16):
    """a [M, K/2] and b [K/2, N] (= w.t() of an [N, K/2] weight) hold two e2m1 codes per byte (low nibble first).
    Scales are e4m3 bytes in the 128x4-tiled layout: byte of (row r, k-block c) sits at
    ((r // 128) * ceil(kb / 4) + c // 4) * 512 + (r % 32) * 16 + ((r % 128) // 32) * 4 + c % 4.   out = alpha * deq(a) @ deq(w).T"""
    grid = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0], device=a.device)

    def dequant(packed, sf):
        rows, kb = packed.shape[0], packed.shape[1] * 2 // block_size
        r = torch.arange(rows, device=packed.device)[:, None]
        c = torch.arange(kb, device=packed.device)[None, :]
        off = ((r // 128) * ((kb + 3) // 4) + c // 4) * 512 + (r % 32) * 16 + ((r % 128) // 32) * 4 + c % 4
        scale = sf.reshape(-1).view(torch.uint8)[off].view(torch.float8_e4m3fn).to(torch.float32).repeat_interleave(block_size, dim=1)
        codes = torch.stack([packed & 0xF, packed >> 4], dim=-1).reshape(rows, -1).long()
        return grid[codes & 7] * torch.where(codes >= 8, -1.0, 1.0) * scale

    out = dequant(a, a_descale) @ dequant(b.t(), b_descale).t()
    if alpha is not None:
        out = out * alpha.to(torch.float32)
    return out.to(torch.bfloat16)


def _mm_fp4_init(*, M=64, N=4096, K=4096, device="cuda", seed=0):
    import flashinfer_b200 as fi

    _, a, w = _ab(M, N, K, device, seed)
    ga, gw = (448.0 * 6.0) / a.float().abs().max(), (448.0 * 6.0) / w.float().abs().max()
    aq, asf = fi.fp4_quantize(a, ga.reshape(1))                  # swizzled scales (the default)
    wq, wsf = fi.fp4_quantize(w, gw.reshape(1))
    return {"a": aq, "b": wq.t(), "a_descale": asf.reshape(-1), "b_descale": wsf.reshape(-1), "alpha": (1.0 / (ga * gw)).reshape(1).to(device),
            "out_dtype": torch.bfloat16}


mm_fp4_trace = TraceTemplate(
    op_type="gemm", name_fmt="mm_fp4_n{N}_k{K}", axes=[Var("M"), Const("N"), Const("K")],
    inputs=[Tensor("a", ("M", "K_half"), "uint8"), Tensor("b", ("K_half", "N"), "uint8"), Tensor("a_descale", ("a_scale_bytes",), "uint8"),
            Tensor("b_descale", ("b_scale_bytes",), "uint8"), Tensor("alpha", ("one",), "float32", optional=True), Scalar("block_size", "int32", optional=True)],
    outputs=[Tensor("out", ("M", "N"), dtype="bfloat16")], reference=_mm_fp4_reference, init=_mm_fp4_init, tags=("gemm", "nvfp4"),
    constraints=("K_half == K / 2", "a_scale_bytes == round_up(M, 128) * round_up(K / 16, 4)", "b_scale_bytes == round_up(N, 128) * round_up(K / 16, 4)",
                 "one == 1"), derive=lambda s: {"K": 2 * s["K_half"]} if "K_half" in s else {},
    description="NVFP4 block-scaled GEMM (tcgen05 kind::mxf4nvf4): e2m1 operands, e4m3 scale per 16 elements, global alpha",
    tolerance="cos", test_sizes={"N": 128, "K": 128})


def _mm_mxfp8_reference(a, b, a_descale, b_descale):
    """a [M, K] e4m3, b [K, N] (= w.t() of an [N, K] e4m3 weight); UE8M0 scale bytes per 32 elements in the 128x4-tiled layout
    (see mm_fp4): value = 2^(byte - 127)."""
    def dequant(x, sf):
        rows, kb = x.shape[0], x.shape[1] // 32
        r = torch.arange(rows, device=x.device)[:, None]
        c = torch.arange(kb, device=x.device)[None, :]
        off = ((r // 128) * ((kb + 3) // 4) + c // 4) * 512 + (r % 32) * 16 + ((r % 128) // 32) * 4 + c % 4
        scale = torch.exp2(sf.reshape(-1).view(torch.uint8)[off].to(torch.float32) - 127.0).repeat_interleave(32, dim=1)
        return x.to(torch.float32) * scale

    return (dequant(a, a_descale) @ dequant(b.t(), b_descale).t()).to(torch.bfloat16)


def _mm_mxfp8_init(*, M=64, N=4096, K=4096, device="cuda", seed=0):
    import flashinfer_b200 as fi

    _, a, w = _ab(M, N, K, device, seed)
    aq, asf = fi.mxfp8_quantize(a)
    wq, wsf = fi.mxfp8_quantize(w)
    return {"a": aq, "b": wq.t(), "a_descale": asf.reshape(-1), "b_descale": wsf.reshape(-1), "out_dtype": torch.bfloat16}


mm_mxfp8_trace = TraceTemplate(
    op_type="gemm", name_fmt="mm_mxfp8_n{N}_k{K}", axes=[Var("M"), Const("N"), Const("K")],
    inputs=[Tensor("a", ("M", "K"), "float8_e4m3fn"), Tensor("b", ("K", "N"), "float8_e4m3fn"), Tensor("a_descale", ("a_scale_bytes",), "uint8"),
            Tensor("b_descale", ("b_scale_bytes",), "uint8")],
    outputs=[Tensor("out", ("M", "N"), dtype="bfloat16")], reference=_mm_mxfp8_reference, init=_mm_mxfp8_init, tags=("gemm", "mxfp8"),
    constraints=("a_scale_bytes == round_up(M, 128) * round_up(K / 32, 4)", "b_scale_bytes == round_up(N, 128) * round_up(K / 32, 4)"),
    description="MXFP8 block-scaled GEMM (tcgen05 kind::mxf8f6f4): e4m3 operands with a power-of-two scale per 32 elements",
    tolerance="cos", test_sizes={"N": 128, "K": 128})
