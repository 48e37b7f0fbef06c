"""Rotary-embedding templates (reference flashinfer/trace/templates/rope.py).  Angle of pair ``j`` at position ``p`` is
``p * theta^(-2j / rotary_dim) / scale`` (Llama-3.1: per-frequency smoothing between scaled and unscaled).  Non-interleaved
pairs are (j, j + rotary_dim/2); interleaved pairs are (2j, 2j+1).  Dimensions past ``rotary_dim`` pass through."""
import torch

from ..template import Const, Scalar, Tensor, TraceTemplate, Var

_AXES = [Var("nnz"), Const("num_q_heads", abbrev="hq"), Const("num_k_heads", abbrev="hk"), Const("head_dim", abbrev="d")]
_Q, _K = ("nnz", "num_q_heads", "head_dim"), ("nnz", "num_k_heads", "head_dim")
_OPTS = [Scalar("rotary_dim", "int32", optional=True), Scalar("interleave", "bool", optional=True),
         Scalar("rope_scale", optional=True), Scalar("rope_theta", optional=True)]
_L31 = [Scalar("low_freq_factor", optional=True), Scalar("high_freq_factor", optional=True),
        Scalar("old_context_len", "int32", optional=True)]
_SIZES = {"num_q_heads": 4, "num_k_heads": 2, "head_dim": 64}


def _rope_pos_ids_reference(q, k, pos_ids, rotary_dim=None, interleave=False, rope_scale=1.0, rope_theta=1e4):
    def rotate(x):
        d = x.shape[-1]
        rd = rotary_dim or d
        inv = 1.0 / (rope_theta ** (torch.arange(0, rd, 2, dtype=torch.float32, device=x.device) / rd)) / rope_scale
        ang = pos_ids.to(torch.float32)[:, None] * inv[None, :]
        cos, sin = torch.cos(ang)[:, None, :], torch.sin(ang)[:, None, :]
        xf = x.to(torch.float32)
        xr = xf[..., :rd]
        if interleave:
            a, b = xr[..., 0::2], xr[..., 1::2]
            rot = torch.stack([a * cos - b * sin, b * cos + a * sin], dim=-1).flatten(-2)
        else:
            a, b = xr[..., : rd // 2], xr[..., rd // 2:]
            rot = torch.cat([a * cos - b * sin, b * cos + a * sin], dim=-1)
        return torch.cat([rot, xf[..., rd:]], dim=-1).to(x.dtype)

    return rotate(q), rotate(k)


def _rope_indptr_reference(q, k, indptr, offsets, rotary_dim=None, interleave=False, rope_scale=1.0, rope_theta=1e4):
    pos = torch.zeros(q.shape[0], dtype=torch.int64, device=q.device)
    for b in range(indptr.numel() - 1):
        s, e = int(indptr[b]), int(indptr[b + 1])
        pos[s:e] = int(offsets[b]) + torch.arange(e - s, device=q.device)

    def rotate(x):
        d = x.shape[-1]
        rd = rotary_dim or d
        inv = 1.0 / (rope_theta ** (torch.arange(0, rd, 2, dtype=torch.float32, device=x.device) / rd)) / rope_scale
        ang = pos.to(torch.float32)[:, None] * inv[None, :]
        cos, sin = torch.cos(ang)[:, None, :], torch.sin(ang)[:, None, :]
        xf = x.to(torch.float32)
        xr = xf[..., :rd]
        if interleave:
            a, b = xr[..., 0::2], xr[..., 1::2]
            rot = torch.stack([a * cos - b * sin, b * cos + a * sin], dim=-1).flatten(-2)
        else:
            a, b = xr[..., : rd // 2], xr[..., rd // 2:]
            rot = torch.cat([a * cos - b * sin, b * cos + a * sin], dim=-1)
        return torch.cat([rot, xf[..., rd:]], dim=-1).to(x.dtype)

    return rotate(q), rotate(k)




def _llama31_rope_pos_ids_reference(q, k, pos_ids, rotary_dim=None, interleave=False, rope_scale=8.0, rope_theta=5e5,
                                    low_freq_factor=1.0, high_freq_factor=4.0, old_context_len=8192):
    import math

    def rotate(x):
        d = x.shape[-1]
        rd = rotary_dim or d
        inv = 1.0 / (rope_theta ** (torch.arange(0, rd, 2, dtype=torch.float32, device=x.device) / rd))
        sa = old_context_len / (2 * math.pi * (high_freq_factor - low_freq_factor))
        sb = -1.0 / (high_freq_factor / low_freq_factor - 1.0)
        smooth = torch.clamp(inv * sa + sb, 0.0, 1.0)
        inv = (1 - smooth) * (inv / rope_scale) + smooth * inv
        ang = pos_ids.to(torch.float32)[:, None] * inv[None, :]
        cos, sin = torch.cos(ang)[:, None, :], torch.sin(ang)[:, None, :]
        xf = x.to(torch.float32)
        xr = xf[..., :rd]
        if interleave:
            a, b = xr[..., 0::2], xr[..., 1::2]
            rot = torch.stack([a * cos - b * sin, b * cos + a * sin], dim=-1).flatten(-2)
        else:
            a, b = xr[..., : rd // 2], xr[..., rd // 2:]
            rot = torch.cat([a * cos - b * sin, b * cos + a * sin], dim=-1)
        return torch.cat([rot, xf[..., rd:]], dim=-1).to(x.dtype)

    return rotate(q), rotate(k)


def _llama31_rope_indptr_reference(q, k, indptr, offsets, rotary_dim=None, interleave=False, rope_scale=8.0, rope_theta=5e5,
                                   low_freq_factor=1.0, high_freq_factor=4.0, old_context_len=8192):
    import math

    pos = torch.zeros(q.shape[0], dtype=torch.int64, device=q.device)
    for b in range(indptr.numel() - 1):
        s, e = int(indptr[b]), int(indptr[b + 1])
        pos[s:e] = int(offsets[b]) + torch.arange(e - s, device=q.device)

    def rotate(x):
        d = x.shape[-1]
        rd = rotary_dim or d
        inv = 1.0 / (rope_theta ** (torch.arange(0, rd, 2, dtype=torch.float32, device=x.device) / rd))
        sa = old_context_len / (2 * math.pi * (high_freq_factor - low_freq_factor))
        sb = -1.0 / (high_freq_factor / low_freq_factor - 1.0)
        smooth = torch.clamp(inv * sa + sb, 0.0, 1.0)
        inv = (1 - smooth) * (inv / rope_scale) + smooth * inv
        ang = pos.to(torch.float32)[:, None] * inv[None, :]
        cos, sin = torch.cos(ang)[:, None, :], torch.sin(ang)[:, None, :]
        xf = x.to(torch.float32)
        xr = xf[..., :rd]
        if interleave:
            a, b = xr[..., 0::2], xr[..., 1::2]
            rot = torch.stack([a * cos - b * sin, b * cos + a * sin], dim=-1).flatten(-2)
        else:
            a, b = xr[..., : rd // 2], xr[..., rd // 2:]
            rot = torch.cat([a * cos - b * sin, b * cos + a * sin], dim=-1)
        return torch.cat([rot, xf[..., rd:]], dim=-1).to(x.dtype)

    return rotate(q), rotate(k)


def _qk(nnz, hq, hk, d, device, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    q = torch.randn(nnz, hq, d, generator=g).to(torch.bfloat16).to(device)
    k = torch.randn(nnz, hk, d, generator=g).to(torch.bfloat16).to(device)
    return g, q, k


def _pos_ids_init(*, nnz=64, num_q_heads=32, num_k_heads=8, head_dim=128, device="cuda", seed=0):
    g, q, k = _qk(nnz, num_q_heads, num_k_heads, head_dim, device, seed)
    pos = torch.randint(0, 4096, (nnz,), generator=g, dtype=torch.int32).to(device)
    return {"q": q, "k": k, "pos_ids": pos}


def _indptr_init(*, nnz=64, num_q_heads=32, num_k_heads=8, head_dim=128, device="cuda", seed=0):
    g, q, k = _qk(nnz, num_q_heads, num_k_heads, head_dim, device, seed)
    cut = max(1, nnz // 3)
    indptr = torch.tensor([0, cut, nnz], dtype=torch.int32, device=device)
    offsets = torch.tensor([7, 1000], dtype=torch.int32, device=device)
    return {"q": q, "k": k, "indptr": indptr, "offsets": offsets}


def _partial_interleaved(init):
    def build(**kw):
        out = init(**kw)
        out.update(rotary_dim=out["q"].shape[-1] // 2, interleave=True)
        return out

    build.__signature__ = __import__("inspect").signature(init)
    return build


def _make(name, ref, init, indptr: bool, inplace: bool, llama31: bool, desc):
    idx = [Tensor("indptr", ("len_indptr",), "int32"), Tensor("offsets", ("batch_size",), "int32")] if indptr else \
        [Tensor("pos_ids", ("nnz",), "int32")]
    outs = [Tensor("q_rope", _Q, dtype_from="q", param="q" if inplace else None),
            Tensor("k_rope", _K, dtype_from="k", param="k" if inplace else None)]
    return TraceTemplate(
        op_type="rope", name_fmt=name + "_hq{num_q_heads}_hk{num_k_heads}_d{head_dim}", axes=_AXES + ([Var("batch_size")] if indptr else []),
        inputs=[Tensor("q", _Q), Tensor("k", _K)] + idx + _OPTS + (_L31 if llama31 else []), outputs=outs, reference=ref, init=init,
        tags=("rope",) + (("inplace",) if inplace else ()) + (("llama31",) if llama31 else ()), description=desc,
        constraints=("len_indptr == batch_size + 1",) if indptr else (), tolerance="bf16", test_sizes=_SIZES)


apply_rope_trace = _make("apply_rope", _rope_indptr_reference, _indptr_init, True, False, False,
                         "RoPE over ragged q / k; positions = offsets[b] + index within segment b")
apply_rope_inplace_trace = _make("apply_rope_inplace", _rope_indptr_reference, _partial_interleaved(_indptr_init), True, True, False,
                                 "In-place RoPE over ragged q / k (exercised with a partial, interleaved rotation)")
apply_rope_pos_ids_trace = _make("apply_rope_pos_ids", _rope_pos_ids_reference, _pos_ids_init, False, False, False,
                                 "RoPE with explicit per-token positions")
apply_rope_pos_ids_inplace_trace = _make("apply_rope_pos_ids_inplace", _rope_pos_ids_reference, _partial_interleaved(_pos_ids_init),
                                         False, True, False, "In-place RoPE with explicit positions")
apply_llama31_rope_trace = _make("apply_llama31_rope", _llama31_rope_indptr_reference, _indptr_init, True, False, True,
                                 "Llama-3.1 frequency-smoothed RoPE over ragged q / k")
apply_llama31_rope_inplace_trace = _make("apply_llama31_rope_inplace", _llama31_rope_indptr_reference, _indptr_init, True, True, True,
                                         "In-place Llama-3.1 RoPE over ragged q / k")
apply_llama31_rope_pos_ids_trace = _make("apply_llama31_rope_pos_ids", _llama31_rope_pos_ids_reference, _pos_ids_init, False, False,
                                         True, "Llama-3.1 RoPE with explicit positions")
apply_llama31_rope_pos_ids_inplace_trace = _make("apply_llama31_rope_pos_ids_inplace", _llama31_rope_pos_ids_reference, _pos_ids_init,
                                                 False, True, True, "In-place Llama-3.1 RoPE with explicit positions")


# ---- cos/sin-cache flavour (vLLM / SGLang call signature: flattened heads)
def _cos_sin_cache_reference(positions, query, key, head_size, cos_sin_cache, is_neox=True):
    rd = cos_sin_cache.shape[-1]
    cs = cos_sin_cache[positions.long()].to(torch.float32)
    cos, sin = cs[:, None, : rd // 2], cs[:, None, rd // 2:]

    def rotate(x):
        xf = x.to(torch.float32).view(x.shape[0], -1, head_size)
        xr = xf[..., :rd]
        if is_neox:
            a, b = xr[..., : rd // 2], xr[..., rd // 2:]
            rot = torch.cat([a * cos - b * sin, b * cos + a * sin], dim=-1)
        else:
            a, b = xr[..., 0::2], xr[..., 1::2]
            rot = torch.stack([a * cos - b * sin, b * cos + a * sin], dim=-1).flatten(-2)
        return torch.cat([rot, xf[..., rd:]], dim=-1).to(x.dtype).view(x.shape)

    return rotate(query), rotate(key)


def _cos_sin_cache_init(*, nnz=64, q_size=4096, k_size=1024, head_size=128, rotary_dim=128, max_position=8192, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    inv = 1.0 / (1e4 ** (torch.arange(0, rotary_dim, 2, dtype=torch.float32) / rotary_dim))
    ang = torch.arange(max_position, dtype=torch.float32)[:, None] * inv[None, :]
    cache = torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1).to(device)
    return {"positions": torch.randint(0, max_position, (nnz,), generator=g).to(device),
            "query": torch.randn(nnz, q_size, generator=g).to(torch.bfloat16).to(device),
            "key": torch.randn(nnz, k_size, generator=g).to(torch.bfloat16).to(device), "head_size": head_size,
            "cos_sin_cache": cache, "is_neox": True}


def _make_cache(name, inplace):
    return TraceTemplate(
        op_type="rope", name_fmt=name + "_q{q_size}_k{k_size}_d{head_size}_r{rotary_dim}",
        axes=[Var("nnz"), Const("q_size"), Const("k_size"), Const("head_size"), Const("rotary_dim"), Var("max_position")],
        inputs=[Tensor("positions", ("nnz",)), Tensor("query", ("nnz", "q_size")), Tensor("key", ("nnz", "k_size")),
                Scalar("head_size", "int32"), Tensor("cos_sin_cache", ("max_position", "rotary_dim"), description="[cos | sin] halves, fp32"),
                Scalar("is_neox", "bool", optional=True)],
        outputs=[Tensor("query_out", ("nnz", "q_size"), dtype_from="query", param="query" if inplace else None),
                 Tensor("key_out", ("nnz", "k_size"), dtype_from="key", param="key" if inplace else None)],
        reference=_cos_sin_cache_reference, init=_cos_sin_cache_init, tags=("rope", "cos_sin_cache") + (("inplace",) if inplace else ()),
        description="RoPE from a precomputed cos/sin table over flattened-head q / k", tolerance="bf16",
        test_sizes={"q_size": 256, "k_size": 128, "head_size": 64, "rotary_dim": 32, "max_position": 128})


apply_rope_with_cos_sin_cache_trace = _make_cache("apply_rope_with_cos_sin_cache", False)
apply_rope_with_cos_sin_cache_inplace_trace = _make_cache("apply_rope_with_cos_sin_cache_inplace", True)


# ---- RoPE + fp8 quantisation of split (rope | nope) heads: the MLA / DeepSeek pre-attention step
def _rope_quantize_fp8_reference(q_rope, k_rope, q_nope, k_nope, cos_sin_cache, pos_ids, is_neox=True, quant_scale_q=1.0, quant_scale_kv=1.0):
    """Rotate the rope slices with the cached cos / sin of each token's position, multiply every slice by its quantisation
    scale and cast to e4m3 (saturating).  k slices may be [nnz, d] (one shared head, MLA) or [nnz, Hk, d]."""
    rd = cos_sin_cache.shape[-1]
    cs = cos_sin_cache[pos_ids.long()].to(torch.float32)
    cos, sin = cs[:, None, : rd // 2], cs[:, None, rd // 2:]

    def rotate(x):
        xf = x.to(torch.float32)
        xf = xf[:, None, :] if x.dim() == 2 else xf
        if is_neox:
            a, b = xf[..., : rd // 2], xf[..., rd // 2: rd]
            rot = torch.cat([a * cos - b * sin, b * cos + a * sin], dim=-1)
        else:
            a, b = xf[..., 0:rd:2], xf[..., 1:rd:2]
            rot = torch.stack([a * cos - b * sin, b * cos + a * sin], dim=-1).flatten(-2)
        rot = torch.cat([rot, xf[..., rd:]], dim=-1)
        return rot[:, 0, :] if x.dim() == 2 else rot

    def quant(x, s):
        return (x.to(torch.float32) * s).clamp(-448.0, 448.0).to(torch.float8_e4m3fn)

    return quant(rotate(q_rope), quant_scale_q), quant(rotate(k_rope), quant_scale_kv), quant(q_nope, quant_scale_q), quant(k_nope, quant_scale_kv)


def _rope_quantize_fp8_init(*, nnz=64, num_heads=128, rope_dim=64, nope_dim=512, max_position=4096, device="cuda", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    inv = 1.0 / (1e4 ** (torch.arange(0, rope_dim, 2, dtype=torch.float32) / rope_dim))
    ang = torch.arange(max_position, dtype=torch.float32)[:, None] * inv[None, :]
    mk = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16).to(device)  # noqa: E731
    return {"q_rope": mk(nnz, num_heads, rope_dim), "k_rope": mk(nnz, rope_dim), "q_nope": mk(nnz, num_heads, nope_dim), "k_nope": mk(nnz, nope_dim),
            "cos_sin_cache": torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1).to(device),
            "pos_ids": torch.randint(0, max_position, (nnz,), generator=g, dtype=torch.int32).to(device), "is_neox": True,
            "quant_scale_q": 0.5, "quant_scale_kv": 0.25}


def _fp8_pairs_compare(got, expected, kwargs):
    for g_, e_ in zip(got, expected):
        assert g_.dtype == torch.float8_e4m3fn and g_.shape == e_.shape
        torch.testing.assert_close(g_.to(torch.float32), e_.to(torch.float32), atol=2.0 ** -6, rtol=0.13)     # one e4m3 step


def _make_rq(name, desc):
    return TraceTemplate(
        op_type="rope", name_fmt=name + "_h{num_heads}_r{rope_dim}_n{nope_dim}",
        axes=[Var("nnz"), Var("max_position"), Const("num_heads", abbrev="h"), Const("rope_dim", abbrev="r"), Const("nope_dim", abbrev="n")],
        inputs=[Tensor("q_rope", ("nnz", "num_heads", "rope_dim")), Tensor("k_rope", ("nnz", "rope_dim")), Tensor("q_nope", ("nnz", "num_heads", "nope_dim")),
                Tensor("k_nope", ("nnz", "nope_dim")), Tensor("cos_sin_cache", ("max_position", "rope_dim"), "float32"), Tensor("pos_ids", ("nnz",), "int32"),
                Scalar("is_neox", "bool", optional=True), Scalar("quant_scale_q", optional=True), Scalar("quant_scale_kv", optional=True)],
        outputs=[Tensor("q_rope_out", ("nnz", "num_heads", "rope_dim"), dtype="float8_e4m3fn"), Tensor("k_rope_out", ("nnz", "rope_dim"), dtype="float8_e4m3fn"),
                 Tensor("q_nope_out", ("nnz", "num_heads", "nope_dim"), dtype="float8_e4m3fn"), Tensor("k_nope_out", ("nnz", "nope_dim"), dtype="float8_e4m3fn")],
        reference=_rope_quantize_fp8_reference, init=_rope_quantize_fp8_init, compare=_fp8_pairs_compare, tags=("rope", "fp8", "mla"),
        description=desc, test_sizes={"num_heads": 4, "rope_dim": 16, "nope_dim": 32, "max_position": 64})


rope_quantize_fp8_trace = _make_rq("rope_quantize_fp8", "RoPE on the rotary slices + fp8 quantisation of rope and no-rope slices of q and k")
mla_rope_quantize_fp8_trace = _make_rq("mla_rope_quantize_fp8", "MLA spelling of rope_quantize_fp8 (shared 2-D k slices)")
