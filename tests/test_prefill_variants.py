"""Attention variants inside the tcgen05 prefill kernel (csrc/attention/prefill_sm100.cu softmax passes): packed / dense custom masks,
ALiBi, and user LogitsTransform / LogitsMask hooks JIT-compiled into the kernel through jit.gen_customize_batch_prefill_module.
Reference parity: tests/attention/test_batch_prefill_kernels.py (custom mask), test_alibi.py, tests/utils/test_jit_example.py."""
import math

import pytest
import torch

import flashinfer_b200 as fi
from flashinfer_b200 import jit, reference

VARIANT_DECL = r"""
struct SoftBiasVariant : VariantDefaults {
  // logits * temp + bias[head] * (kv position - query position);   every 7th key (except key 0) is masked out
  static __device__ __forceinline__ float LogitsTransform(const VariantCtx& ctx, float logits, int kv_idx) {
    return logits * temp + bias[ctx.qo_head_idx] * float(kv_idx - (ctx.kv_len - ctx.qo_len + ctx.qo_idx));
  }
  REGISTER_LOGITS_MASK(params, batch_idx, qo_idx, kv_idx, qo_head_idx, kv_head_idx, { return (kv_idx % 7) != 3 || kv_idx == 0; })
};
"""
JIT_ARGS = ["test_softbias", torch.bfloat16, torch.bfloat16, torch.bfloat16, torch.int32, 128, 128, ["bias"], ["float"], ["temp"],
            ["double"], "SoftBiasVariant", VARIANT_DECL]


def test_variant_module_builds():
    """CPU-side: the generator writes the header and nvcc cross-compiles the private prefill module for sm_100a (the .so stays
    in-tree next to the other native modules, so a GPU box loads it without compiling)."""
    if not jit.have_nvcc():
        pytest.skip("nvcc not available")
    spec = jit.gen_customize_batch_prefill_module("auto", *JIT_ARGS)
    jit.build_module(spec)
    assert spec.so_path.exists() and spec.is_fresh()
    assert spec.additional_tensor_names == ("bias",) and spec.additional_scalar_names == ("temp",)
    with pytest.raises(ValueError):
        jit.gen_customize_batch_prefill_module("auto", "bad", None, None, None, None, 128, 128, ["a"] * 9, ["float"] * 9, [], [], "V", "")


def test_variant_argument_names_do_not_leak_into_the_kernel():
    """Additional tensors / scalars are visible to the hooks as macros; they are undefined again after the variant struct, so a
    user scalar called like one of the kernel's own locals (``alpha``, ``l``) compiles."""
    if not jit.have_nvcc():
        pytest.skip("nvcc not available")
    decl = ("struct Clash : VariantDefaults { static __device__ __forceinline__ float LogitsTransform(const VariantCtx& ctx, float logits, "
            "int kv_idx) { return logits * alpha + l[kv_idx]; } };")
    spec = jit.gen_customize_batch_prefill_module("auto", "unit_name_clash", None, None, None, None, 128, 128, ["l"], ["float"], ["alpha"],
                                                  ["double"], "Clash", decl)
    try:
        jit.build_module(spec)
        assert spec.is_fresh()
    finally:
        for p in (spec.so_path, spec.hash_path):
            if p.exists():
                p.unlink()


def _no_generic(monkeypatch):
    """The SIMT catch-all kernel must not serve these configurations any more."""
    from flashinfer_b200.attention import generic

    def boom(*a, **k):
        raise AssertionError("generic (CUDA-core) attention kernel was used")

    monkeypatch.setattr(generic, "run", boom)


def _ragged_case(seed, lens, hq, hkv, d=128, dtype=torch.bfloat16):
    torch.manual_seed(seed)
    qo = torch.tensor([0] + [a for a, _ in lens]).cumsum(0).int()
    kv = torch.tensor([0] + [b for _, b in lens]).cumsum(0).int()
    q = torch.randn(int(qo[-1]), hq, d, device="cuda", dtype=dtype)
    k = torch.randn(int(kv[-1]), hkv, d, device="cuda", dtype=dtype)
    v = torch.randn(int(kv[-1]), hkv, d, device="cuda", dtype=dtype)
    return qo, kv, q, k, v


@pytest.mark.gpu
@pytest.mark.parametrize("packed", [False, True])
def test_custom_mask_ragged_tcgen05(monkeypatch, packed):
    _no_generic(monkeypatch)
    lens = [(37, 90), (300, 300), (1, 513), (129, 140)]
    qo, kv, q, k, v = _ragged_case(0, lens, 8, 2)
    masks = []
    for ql, kl in lens:
        m = torch.rand(ql, kl, device="cuda") > 0.35
        m[:, 0] = True
        masks.append(m)
    masks[1][5, :] = False  # a fully masked row: output 0, lse -inf
    flat = torch.cat([m.flatten() for m in masks])
    w = fi.BatchPrefillWithRaggedKVCacheWrapper(torch.empty(32 << 20, dtype=torch.uint8, device="cuda"))
    # packed form = segment_packbits: every request's q_len x kv_len bits start on a byte boundary (3330 / 513 bits here are not multiples of 8)
    bits = torch.tensor([0] + [a * b for a, b in lens]).cumsum(0).int()
    kw = {"packed_custom_mask": fi.segment_packbits(flat.cpu(), bits, bitorder="little")[0].cuda()} if packed else {"custom_mask": flat}
    w.plan(qo, kv, 8, 2, 128, causal=True, q_data_type=torch.bfloat16, **kw)  # causal is overridden by the custom mask
    out, lse = w.run(q, k, v, return_lse=True)
    for b, (ql, kl) in enumerate(lens):
        qs, ks = int(qo[b]), int(kv[b])
        ref, lref = reference.attention_ref(q[qs:qs + ql], k[ks:ks + kl], v[ks:ks + kl], False, 1 / math.sqrt(128), custom_mask=masks[b])
        ref = torch.nan_to_num(ref.float(), nan=0.0)
        assert (out[qs:qs + ql].float() - ref).abs().max() < 2e-2
        fin = torch.isfinite(lref)
        assert (lse[qs:qs + ql][fin] - lref[fin]).abs().max() < 2e-2
        assert torch.isinf(lse[qs:qs + ql][~fin]).all()


@pytest.mark.gpu
def test_custom_mask_paged_and_single_tcgen05(monkeypatch):
    _no_generic(monkeypatch)
    torch.manual_seed(3)
    hq, hkv, d, ps = 4, 4, 128, 8
    qo_len, kv_len = 37, 90
    mask = torch.rand(qo_len, kv_len, device="cuda") > 0.4
    mask[:, 0] = True
    q = torch.randn(qo_len, hq, d, device="cuda", dtype=torch.float16)
    k = torch.randn(kv_len, hkv, d, device="cuda", dtype=torch.float16)
    v = torch.randn(kv_len, hkv, d, device="cuda", dtype=torch.float16)
    ref, _ = reference.attention_ref(q, k, v, False, 1 / math.sqrt(d), custom_mask=mask)
    out = fi.single_prefill_with_kv_cache(q, k, v, custom_mask=mask)
    assert (out.float() - ref.float()).abs().max() < 2e-2
    n_pages = (kv_len + ps - 1) // ps
    kc = torch.zeros(n_pages, ps, hkv, d, device="cuda", dtype=torch.float16)
    vc = torch.zeros_like(kc)
    kc.view(-1, hkv, d)[:kv_len] = k
    vc.view(-1, hkv, d)[:kv_len] = v
    w = fi.BatchPrefillWithPagedKVCacheWrapper(torch.empty(32 << 20, dtype=torch.uint8, device="cuda"))
    w.plan(torch.tensor([0, qo_len], dtype=torch.int32), torch.tensor([0, n_pages], dtype=torch.int32),
           torch.arange(n_pages, dtype=torch.int32), torch.tensor([(kv_len - 1) % ps + 1], dtype=torch.int32), hq, hkv, d, ps,
           custom_mask=mask.flatten(), q_data_type=torch.float16)
    out2 = w.run(q, (kc, vc))
    assert (out2.float() - ref.float()).abs().max() < 2e-2


@pytest.mark.gpu
@pytest.mark.parametrize("causal", [False, True])
def test_alibi_tcgen05(monkeypatch, causal):
    _no_generic(monkeypatch)
    lens = [(64, 200), (257, 257), (5, 1000)]
    hq, hkv = 12, 4  # non power-of-two head count: both slope series
    qo, kv, q, k, v = _ragged_case(1, lens, hq, hkv)
    w = fi.BatchPrefillWithRaggedKVCacheWrapper(torch.empty(32 << 20, dtype=torch.uint8, device="cuda"))
    w.plan(qo, kv, hq, hkv, 128, causal=causal, pos_encoding_mode="ALIBI", logits_soft_cap=0.0, q_data_type=torch.bfloat16)
    out = w.run(q, k, v)
    slopes = fi.utils.get_alibi_slopes(hq, q.device)
    for b, (ql, kl) in enumerate(lens):
        qs, ks = int(qo[b]), int(kv[b])
        ref, _ = reference.attention_ref(q[qs:qs + ql], k[ks:ks + kl], v[ks:ks + kl], causal, 1 / math.sqrt(128), alibi_slopes=slopes)
        assert (out[qs:qs + ql].float() - ref.float()).abs().max() < 2e-2
    o1 = fi.single_prefill_with_kv_cache(q[:64], k[:200], v[:200], causal=causal, pos_encoding_mode="ALIBI")
    ref, _ = reference.attention_ref(q[:64], k[:200], v[:200], causal, 1 / math.sqrt(128), alibi_slopes=slopes)
    assert (o1.float() - ref.float()).abs().max() < 2e-2


@pytest.mark.gpu
def test_user_variant_jit_tcgen05(monkeypatch):
    """LogitsTransform + LogitsMask supplied as C++ and compiled into the tensor-core kernel (reference tests/utils/test_jit_example.py)."""
    _no_generic(monkeypatch)
    lens = [(100, 260), (33, 33)]
    hq, hkv = 8, 2
    qo, kv, q, k, v = _ragged_case(2, lens, hq, hkv)
    bias = (torch.rand(hq, device="cuda") * 0.02).float()
    temp = 0.7
    w = fi.BatchPrefillWithRaggedKVCacheWrapper(torch.empty(32 << 20, dtype=torch.uint8, device="cuda"), jit_args=JIT_ARGS)
    w.plan(qo, kv, hq, hkv, 128, causal=True, q_data_type=torch.bfloat16)
    out, lse = w.run(q, k, v, bias, temp, return_lse=True)
    sm = 1 / math.sqrt(128)
    for b, (ql, kl) in enumerate(lens):
        qs, ks = int(qo[b]), int(kv[b])
        qq = q[qs:qs + ql].float().transpose(0, 1)                                     # [hq, ql, d]
        kk = k[ks:ks + kl].float().transpose(0, 1).repeat_interleave(hq // hkv, 0)     # [hq, kl, d]
        vv = v[ks:ks + kl].float().transpose(0, 1).repeat_interleave(hq // hkv, 0)
        logits = qq @ kk.transpose(1, 2) * sm
        qpos = (kl - ql + torch.arange(ql, device="cuda"))[:, None]
        kpos = torch.arange(kl, device="cuda")[None, :]
        logits = logits * temp + bias[:, None, None] * (kpos - qpos)[None].float()
        keep = (kpos <= qpos) & ((kpos % 7 != 3) | (kpos == 0))
        logits = logits.masked_fill(~keep[None], float("-inf"))
        ref = (torch.softmax(logits, -1) @ vv).transpose(0, 1)
        assert (out[qs:qs + ql].float() - ref).abs().max() < 2e-2
        lref = (torch.logsumexp(logits, -1) * math.log2(math.e)).transpose(0, 1)
        assert (lse[qs:qs + ql] - lref).abs().max() < 2e-2
    with pytest.raises(ValueError):
        w.run(q, k, v, bias)  # the scalar is missing


@pytest.mark.gpu
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_head_dim_64_tcgen05(monkeypatch, causal, dtype):
    """head_dim 64 (qk = vo) runs on the tcgen05 prefill kernel (128 x 64 PV tiles), ragged and paged, with soft-cap / window."""
    _no_generic(monkeypatch)
    lens = [(200, 200), (1, 777), (300, 1200), (129, 129)]
    hq, hkv, d = 8, 2, 64
    qo, kv, q, k, v = _ragged_case(7, lens, hq, hkv, d=d, dtype=dtype)
    w = fi.BatchPrefillWithRaggedKVCacheWrapper(torch.empty(32 << 20, dtype=torch.uint8, device="cuda"))
    w.plan(qo, kv, hq, hkv, d, causal=causal, q_data_type=dtype)
    out, lse = w.run(q, k, v, return_lse=True)
    for b, (ql, kl) in enumerate(lens):
        qs, ks = int(qo[b]), int(kv[b])
        ref, lref = reference.attention_ref(q[qs:qs + ql], k[ks:ks + kl], v[ks:ks + kl], causal, 1 / math.sqrt(d))
        assert (out[qs:qs + ql].float() - ref.float()).abs().max() < 2e-2
        assert (lse[qs:qs + ql] - lref).abs().max() < 2e-2
    # paged, page_size 16, soft-cap + sliding window
    ps, kl, ql = 16, 1000, 260
    n_pages = (kl + ps - 1) // ps
    perm = torch.randperm(n_pages + 5)[:n_pages].int()
    kc = torch.zeros(n_pages + 5, ps, hkv, d, device="cuda", dtype=dtype)
    vc = torch.zeros_like(kc)
    kk, vv = torch.randn(kl, hkv, d, device="cuda").to(dtype), torch.randn(kl, hkv, d, device="cuda").to(dtype)
    for i in range(n_pages):
        n = min(ps, kl - i * ps)
        kc[perm[i], :n] = kk[i * ps:i * ps + n]
        vc[perm[i], :n] = vv[i * ps:i * ps + n]
    qq = torch.randn(ql, hq, d, device="cuda").to(dtype)
    wp = fi.BatchPrefillWithPagedKVCacheWrapper(torch.empty(32 << 20, dtype=torch.uint8, device="cuda"))
    wp.plan(torch.tensor([0, ql], dtype=torch.int32), torch.tensor([0, n_pages], dtype=torch.int32), perm,
            torch.tensor([(kl - 1) % ps + 1], dtype=torch.int32), hq, hkv, d, ps, causal=True, window_left=300, logits_soft_cap=20.0,
            q_data_type=dtype)
    o2 = wp.run(qq, (kc, vc))
    ref, _ = reference.attention_ref(qq, kk, vv, True, 1 / math.sqrt(d), 20.0, 300)
    assert (o2.float() - ref.float()).abs().max() < 2e-2
    o3 = fi.single_prefill_with_kv_cache(qq, kk, vv, causal=causal)
    ref, _ = reference.attention_ref(qq, kk, vv, causal, 1 / math.sqrt(d))
    assert (o3.float() - ref.float()).abs().max() < 2e-2


@pytest.mark.gpu
def test_attention_sinks_in_kernel(monkeypatch):
    """`sinks=` joins the softmax denominator inside the tcgen05 kernels (decode: unsplit segments, in-kernel split-KV merge;
    paged prefill epilogue) - the torch post-processing of (o, lse) must not run."""
    _no_generic(monkeypatch)
    from flashinfer_b200.attention import _core

    def boom(*a, **k):
        raise AssertionError("attention sink applied post hoc in torch")

    monkeypatch.setattr(_core, "apply_attention_sink", boom)
    torch.manual_seed(11)
    hq, hkv, d, ps = 8, 2, 128, 16
    sinks = torch.randn(hq, device="cuda") * 2
    # ---- decode: a long request (split over CTAs, merged in-kernel) and short ones (unsplit)
    kv_lens = [9000, 17, 300, 1]
    n_pages = [(l + ps - 1) // ps for l in kv_lens]
    indptr = torch.tensor([0] + list(torch.tensor(n_pages).cumsum(0)), dtype=torch.int32)
    indices = torch.randperm(sum(n_pages)).int()
    last = torch.tensor([(l - 1) % ps + 1 for l in kv_lens], dtype=torch.int32)
    kc = torch.randn(sum(n_pages), ps, hkv, d, device="cuda").bfloat16()
    vc = torch.randn(sum(n_pages), ps, hkv, d, device="cuda").bfloat16()
    q = torch.randn(len(kv_lens), hq, d, device="cuda").bfloat16()
    w = fi.BatchDecodeWithPagedKVCacheWrapper(torch.empty(64 << 20, dtype=torch.uint8, device="cuda"))
    w.plan(indptr, indices, last, hq, hkv, d, ps, q_data_type=torch.bfloat16)
    out, lse = w.run(q, (kc, vc), sinks=sinks, return_lse=True)
    for b, L in enumerate(kv_lens):
        pages = indices[int(indptr[b]):int(indptr[b + 1])].long().cuda()
        k = kc[pages].reshape(-1, hkv, d)[:L]
        v = vc[pages].reshape(-1, hkv, d)[:L]
        ref, lref = reference.attention_ref(q[b:b + 1], k, v, False, 1 / math.sqrt(d), sinks=sinks)
        assert (out[b:b + 1].float() - ref.float()).abs().max() < 2e-2
        assert (lse[b:b + 1] - lref).abs().max() < 2e-2
    # ---- paged prefill (causal) with sinks
    ql, kl = 200, 700
    npg = (kl + ps - 1) // ps
    perm = torch.randperm(npg).int()
    qq = torch.randn(ql, hq, d, device="cuda").bfloat16()
    wp = fi.BatchPrefillWithPagedKVCacheWrapper(torch.empty(32 << 20, dtype=torch.uint8, device="cuda"))
    wp.plan(torch.tensor([0, ql], dtype=torch.int32), torch.tensor([0, npg], dtype=torch.int32), perm,
            torch.tensor([(kl - 1) % ps + 1], dtype=torch.int32), hq, hkv, d, ps, causal=True, q_data_type=torch.bfloat16)
    kcp = torch.randn(npg, ps, hkv, d, device="cuda").bfloat16()
    vcp = torch.randn(npg, ps, hkv, d, device="cuda").bfloat16()
    o2, l2 = wp.run(qq, (kcp, vcp), sinks=sinks, return_lse=True)
    k = kcp[perm.long().cuda()].reshape(-1, hkv, d)[:kl]
    v = vcp[perm.long().cuda()].reshape(-1, hkv, d)[:kl]
    ref, lref = reference.attention_ref(qq, k, v, True, 1 / math.sqrt(d), sinks=sinks)
    assert (o2.float() - ref.float()).abs().max() < 2e-2
    assert (l2 - lref).abs().max() < 2e-2


def test_packed_custom_mask_is_segment_packed_cpu():
    """`packed_custom_mask` follows the reference format (segment_packbits: byte-aligned per request), on the wrappers' CPU path."""
    torch.manual_seed(0)
    lens = [(5, 7), (3, 11), (8, 8)]            # 35 / 33 / 64 bits: the first two segments end inside a byte
    hq, hkv, d = 2, 1, 16
    qo = torch.tensor([0] + [a for a, _ in lens]).cumsum(0).int()
    kv = torch.tensor([0] + [b for _, b in lens]).cumsum(0).int()
    q, k, v = torch.randn(int(qo[-1]), hq, d), torch.randn(int(kv[-1]), hkv, d), torch.randn(int(kv[-1]), hkv, d)
    masks = [torch.rand(a, b) > 0.4 for a, b in lens]
    for m in masks:
        m[:, 0] = True
    flat = torch.cat([m.flatten() for m in masks])
    bits = torch.tensor([0] + [a * b for a, b in lens]).cumsum(0).int()
    packed, _ = fi.segment_packbits(flat, bits, bitorder="little")
    assert packed.numel() == 5 + 5 + 8
    w = fi.BatchPrefillWithRaggedKVCacheWrapper(torch.empty(1 << 20, dtype=torch.uint8))
    w.plan(qo, kv, hq, hkv, d, packed_custom_mask=packed, q_data_type=torch.float32)
    out = w.run(q, k, v)
    for b, (ql, kl) in enumerate(lens):
        ref, _ = reference.attention_ref(q[int(qo[b]):int(qo[b + 1])], k[int(kv[b]):int(kv[b + 1])], v[int(kv[b]):int(kv[b + 1])], False,
                                         1 / math.sqrt(d), custom_mask=masks[b])
        assert (out[int(qo[b]):int(qo[b + 1])] - ref).abs().max() < 1e-4
