"""CPU-path tests of the elementwise ops (the fp32 oracles double as the CPU implementation)."""
import math

import pytest
import torch

import flashinfer_b200 as fi
from flashinfer_b200 import activation, cascade, norm, page, reference, rope


def test_single_prefill_plumbing_cpu():
    # BASELINE.json config #1: single_prefill_with_kv_cache fp32 on CPU, 1 head, seqlen 128
    q, k, v = torch.randn(128, 1, 64), torch.randn(128, 1, 64), torch.randn(128, 1, 64)
    o, lse = fi.single_prefill_with_kv_cache(q, k, v, causal=True, return_lse=True)
    p = torch.softmax((q[:, 0] @ k[:, 0].T / 8.0).masked_fill(~torch.tril(torch.ones(128, 128, dtype=torch.bool)), -1e30), -1)
    torch.testing.assert_close(o[:, 0], p @ v[:, 0], rtol=1e-4, atol=1e-4)
    assert lse.shape == (128, 1)


def test_norms_cpu():
    x, r, w = torch.randn(7, 256), torch.randn(7, 256), torch.randn(256)
    y = norm.rmsnorm(x, w)
    torch.testing.assert_close(y, x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6) * w)
    x2, r2 = x.clone(), r.clone()
    norm.fused_add_rmsnorm(x2, r2, w)
    torch.testing.assert_close(r2, x + r)
    torch.testing.assert_close(x2, norm.rmsnorm(x + r, w))
    torch.testing.assert_close(norm.gemma_rmsnorm(x, w), norm.rmsnorm(x, w + 1))
    g, b = torch.randn(256), torch.randn(256)
    torch.testing.assert_close(norm.layernorm(x, g, b), torch.nn.functional.layer_norm(x, (256,), g, b, 1e-6))


def test_activation_cpu():
    x = torch.randn(5, 64)
    torch.testing.assert_close(activation.silu_and_mul(x), torch.nn.functional.silu(x[:, :32]) * x[:, 32:])
    torch.testing.assert_close(activation.gelu_tanh_and_mul(x), torch.nn.functional.gelu(x[:, :32], approximate="tanh") * x[:, 32:])


def test_merge_state_matches_full_attention():
    q, k, v = torch.randn(4, 2, 32), torch.randn(100, 2, 32), torch.randn(100, 2, 32)
    o, s = reference.attention_ref(q, k, v)
    oa, sa = reference.attention_ref(q, k[:37], v[:37])
    ob, sb = reference.attention_ref(q, k[37:], v[37:])
    om, sm = cascade.merge_state(oa, sa, ob, sb)
    torch.testing.assert_close(om, o, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(sm, s, rtol=1e-4, atol=1e-5)
    vs = torch.stack([oa, ob], 1)
    ss = torch.stack([sa, sb], 1)
    o2, s2 = cascade.merge_states(vs, ss)
    torch.testing.assert_close(o2, o, rtol=1e-4, atol=1e-5)
    oa2, sa2 = oa.clone(), sa.clone()
    cascade.merge_state_in_place(oa2, sa2, ob, sb)
    torch.testing.assert_close(oa2, o, rtol=1e-4, atol=1e-5)


def test_rope_cpu_variants_agree():
    nnz, H, D = 10, 2, 64
    q, k = torch.randn(nnz, H, D), torch.randn(nnz, H, D)
    indptr = torch.tensor([0, 4, 10], dtype=torch.int32)
    offsets = torch.tensor([3, 100], dtype=torch.int32)
    pos = torch.tensor([3, 4, 5, 6, 100, 101, 102, 103, 104, 105], dtype=torch.int32)
    a = rope.apply_rope(q, k, indptr, offsets)
    b = rope.apply_rope_pos_ids(q, k, pos)
    torch.testing.assert_close(a[0], b[0])
    torch.testing.assert_close(a[1], b[1])
    # rotation preserves norms
    torch.testing.assert_close(a[0].norm(dim=-1), q.norm(dim=-1), rtol=1e-4, atol=1e-4)
    # cos/sin cache variant equals on-the-fly
    inv = 1.0 / (1e4 ** (torch.arange(0, D, 2).float() / D))
    ang = torch.arange(200).float()[:, None] * inv[None]
    cache = torch.cat([ang.cos(), ang.sin()], -1)
    qc, kc = rope.apply_rope_with_cos_sin_cache(pos, q.reshape(nnz, -1), k.reshape(nnz, -1), D, cache, is_neox=True)
    torch.testing.assert_close(qc.view(nnz, H, D), b[0], rtol=1e-4, atol=1e-4)


def test_page_append_and_positions_cpu():
    from helpers import make_paged

    kv_lens = [20, 33]
    indptr, indices, last, kc, vc = make_paged(kv_lens, 2, 8, 16)
    append_indptr = torch.tensor([0, 5, 8], dtype=torch.int32)
    seq = torch.tensor(kv_lens, dtype=torch.int32)
    bi, pos = page.get_batch_indices_positions(append_indptr, seq, 8)
    assert bi.tolist() == [0] * 5 + [1] * 3
    assert pos.tolist() == [15, 16, 17, 18, 19, 30, 31, 32]
    k, v = torch.randn(8, 2, 8), torch.randn(8, 2, 8)
    page.append_paged_kv_cache(k, v, bi, pos, (kc, vc), indices, indptr, last)
    kk, vv = reference.gather_paged_kv(kc, vc, indices, indptr, last, 1)
    torch.testing.assert_close(kk[30:33], k[5:])
    torch.testing.assert_close(vv[30:33], v[5:])
    assert page.get_seq_lens(indptr, last, 16).tolist() == kv_lens


def test_fp4_oracle_rounds_ties_to_even():
    """All 7 midpoints of the e2m1 grid round to the even code (cvt.rn semantics) - ADVICE r1."""
    from flashinfer_b200.quantization.fp4 import E2M1_VALUES, fp4_quantize

    mids = [0.25, 0.75, 1.25, 1.75, 2.5, 3.5, 5.0]
    x = torch.tensor([mids + [6.0] + [0.0] * 8], dtype=torch.bfloat16)  # amax 6 -> block scale 1
    q, _ = fp4_quantize(x, torch.tensor([1.0]), 16, False, False)
    codes = torch.stack([q[0] & 0xF, q[0] >> 4], -1).reshape(-1)[:7]
    assert [E2M1_VALUES[int(c)] for c in codes] == [0.0, 1.0, 1.0, 2.0, 2.0, 4.0, 4.0]
    qn, _ = fp4_quantize(-x, torch.tensor([1.0]), 16, False, False)
    assert torch.equal(qn[0, :4] & 0x77, q[0, :4] & 0x77)  # magnitudes symmetric


def test_mm_fp4_accepts_transposed_scale_view():
    """The reference idiom mm_fp4(a, b.T, a_sf, b_sf.T, alpha) (tests/gemm/test_mm_fp4.py) passes a transposed VIEW of the
    swizzled weight scales: the result must match the call with the plain tensor."""
    torch.manual_seed(0)
    m, n, k = 48, 256, 128
    a = torch.randn(m, k, dtype=torch.bfloat16)
    w = torch.randn(n, k, dtype=torch.bfloat16)
    ga = (448.0 * 6.0) / a.float().abs().amax()
    gw = (448.0 * 6.0) / w.float().abs().amax()
    aq, asf = fi.nvfp4_quantize(a, ga.reshape(1))
    wq, wsf = fi.nvfp4_quantize(w, gw.reshape(1))
    assert wsf.dim() == 2
    alpha = 1.0 / (ga * gw)
    plain = fi.mm_fp4(aq, wq.T, asf, wsf, alpha, torch.bfloat16)
    viewed = fi.mm_fp4(aq, wq.T, asf, wsf.T, alpha, torch.bfloat16)
    assert torch.equal(plain, viewed)
    ref = a.float() @ w.float().t()
    cos = torch.nn.functional.cosine_similarity(plain.float().reshape(-1), ref.reshape(-1), dim=0)
    assert cos > 0.97


def test_fused_rmsnorm_silu_nvfp4_output_and_fp4_8x4_layout():
    import flashinfer_b200 as fi
    from flashinfer_b200.quantization.fp4 import SfLayout, _index_8x4, e2m1_and_ufp8sf_scale_to_float, nvfp4_quantize

    torch.manual_seed(0)
    x, w = torch.randn(20, 128).bfloat16(), (1 + 0.1 * torch.randn(128)).bfloat16()
    q, sf = fi.norm.fused_rmsnorm_silu(x, w, 1e-6, out=torch.empty(20, 64, dtype=torch.uint8))
    ref = fi.norm.fused_rmsnorm_silu(x, w, 1e-6).float()
    d = e2m1_and_ufp8sf_scale_to_float(q.view(torch.uint8), sf.view(torch.uint8), None, 16, 1, False)
    assert sf.shape == (20, 8) and torch.nn.functional.cosine_similarity(d.flatten(), ref.flatten(), dim=0) > 0.99
    gs = torch.tensor([100.0])
    q8, sf8 = nvfp4_quantize(x, gs, sfLayout=SfLayout.layout_8x4)
    ql, sfl = nvfp4_quantize(x, gs, sfLayout=SfLayout.layout_linear)
    assert torch.equal(q8, ql) and sf8.numel() == 24 * 8 and torch.equal(sf8.reshape(-1)[_index_8x4(20, 8)], sfl.reshape(-1))


def test_mm_fp4_accepts_8x4_activation_scales():
    """use_8x4_sf_layout (reference mm_fp4 / mm_mxfp8): activation scales in 8x4 tiles give the same product as the 128x4 layout."""
    import flashinfer_b200 as fi
    from flashinfer_b200.quantization.fp4 import SfLayout, nvfp4_quantize

    torch.manual_seed(0)
    m, n, k = 24, 64, 128
    a, w = torch.randn(m, k).bfloat16(), torch.randn(n, k).bfloat16()
    ga, gw = torch.tensor([448 * 6 / float(a.abs().max())]), torch.tensor([448 * 6 / float(w.abs().max())])
    aq8, asf8 = nvfp4_quantize(a, ga, sfLayout=SfLayout.layout_8x4)
    aq, asf = nvfp4_quantize(a, ga)
    wq, wsf = nvfp4_quantize(w, gw)
    alpha = 1 / (ga * gw)
    o8 = fi.mm_fp4(aq8, wq.t(), asf8, wsf.t(), alpha, torch.bfloat16, use_8x4_sf_layout=True)
    assert torch.equal(o8, fi.mm_fp4(aq, wq.t(), asf, wsf.t(), alpha, torch.bfloat16))
    assert torch.nn.functional.cosine_similarity(o8.float().flatten(), (a.float() @ w.float().t()).flatten(), dim=0) > 0.98


def test_logits_processor_modules_and_validators():
    """The pipeline is split like the reference (types / op / operators / processors / legalization / fusion_rules / compiler /
    validators); validators reject pipelines that cannot be lowered; custom checks and custom fusion rules plug in."""
    from flashinfer_b200.logits_processor import LogitsPipe, Sample, Softmax, Temperature, TopK
    from flashinfer_b200.logits_processor.fusion_rules import DEFAULT_RULES, FusionRule
    from flashinfer_b200.logits_processor.legalization import legalize_processors
    from flashinfer_b200.logits_processor.operators import TempSoftmaxOp
    from flashinfer_b200.logits_processor.types import LegalizationError, TensorType
    from flashinfer_b200.logits_processor.validators import validate_pipeline

    with pytest.raises(LegalizationError):
        LogitsPipe([Sample(), Softmax()])
    with pytest.raises(LegalizationError):
        validate_pipeline([Softmax(), Sample(), Sample()])
    with pytest.raises(RuntimeError):
        LogitsPipe([Softmax()], custom_validity_checks=[lambda ps: (_ for _ in ()).throw(RuntimeError("nope"))])
    ops = legalize_processors([Temperature(), Softmax(), TopK()], TensorType.LOGITS)
    assert [o.name for o in ops] == ["temperature", "softmax", "topk_renorm_probs"]
    pipe = LogitsPipe([Temperature(), Softmax(), TopK()])
    assert [o.name for o in pipe.compiled_ops] == ["temperature_softmax", "topk_renorm_probs"] and isinstance(pipe.compiled_ops[0], TempSoftmaxOp)
    extra = FusionRule(("temperature_softmax", "topk_renorm_probs"), lambda ops: ops[0], priority=0)   # user rule: runs after the defaults
    assert len(LogitsPipe([Temperature(), Softmax(), TopK()], custom_fusion_rules=[extra]).compiled_ops) == 1 and len(DEFAULT_RULES) == 6
    x = torch.randn(3, 50)
    probs = pipe(x, temperature=0.7, top_k=4)
    assert torch.allclose(probs.sum(-1), torch.ones(3), atol=1e-5) and int((probs > 0).sum(-1).max()) <= 4


def test_entry_points_refuse_arguments_they_do_not_implement():
    """Signature-parity arguments that would change the result must fail loudly instead of being dropped (ADVICE r1 class of
    bug: silently ignored parameters)."""
    from flashinfer_b200.utils import reject_unsupported

    reject_unsupported("f", a=None, b=False, c=(True, True), d=(None, 3))
    with pytest.raises(NotImplementedError, match="mask', 'c"):
        reject_unsupported("f", mask=torch.zeros(1), c=(False, True))
    q = torch.randn(2, 4, 64, dtype=torch.bfloat16)
    kc = torch.randn(3, 2, 8, 64, dtype=torch.bfloat16)
    ws = torch.empty(1 << 20, dtype=torch.uint8)
    table, lens = torch.tensor([[0, 1], [2, 0]], dtype=torch.int32), torch.tensor([12, 5], dtype=torch.int32)
    with pytest.raises(NotImplementedError):
        fi.decode.trtllm_batch_decode_with_kv_cache(q, (kc, kc), ws, table, lens, 12, mask=torch.zeros(1))
    with pytest.raises(NotImplementedError):
        fi.decode.trtllm_batch_decode_with_kv_cache(q, (kc, kc), ws, table, lens, 12, o_sf_scale=1.0)
    o16 = fi.decode.trtllm_batch_decode_with_kv_cache(q, (kc, kc), ws, table, lens, 12, bmm1_scale=0.125, out_dtype=torch.float16)
    assert o16.dtype == torch.float16
    o, l = fi.decode.trtllm_batch_decode_with_kv_cache(q, (kc, kc), ws, table, lens, 12, bmm1_scale=0.125, return_lse=True,
                                                       lse=torch.empty(2, 4))
    assert l.shape == (2, 4) and torch.allclose(o.float(), o16.float(), atol=2e-2)
    w = fi.BatchPrefillWithRaggedKVCacheWrapper(ws)
    ind = torch.tensor([0, 2], dtype=torch.int32)
    with pytest.raises(NotImplementedError):
        w.plan(ind, ind, 4, 2, 64, q_data_type=torch.bfloat16, prefix_len_ptr=torch.zeros(1, dtype=torch.int32))
    w.plan(ind, ind, 4, 2, 64, q_data_type=torch.bfloat16, o_data_type=torch.float8_e4m3fn)      # a conversion of the kernel's output
    assert w.run(torch.randn(2, 4, 64).bfloat16(), torch.randn(2, 2, 64).bfloat16(), torch.randn(2, 2, 64).bfloat16()).dtype == torch.float8_e4m3fn
    w.plan(ind, ind, 4, 2, 64, q_data_type=torch.bfloat16, o_data_type="bfloat16")
    with pytest.raises(ValueError, match="NaN"):
        fi.sampling.top_k_sampling_from_probs(torch.tensor([[0.5, float("nan"), 0.5]]), 2, check_nan=True)
    with pytest.raises(NotImplementedError):
        fi.quantization.fp4.nvfp4_quantize(torch.randn(4, 64), torch.ones(1), per_token_activation=True)


def test_xqa_batch_decode_default_layout_is_nhd():
    """The reference's XQA batch entry point defaults to NHD pages (its trtllm sibling to HND)."""
    torch.manual_seed(0)
    q = torch.randn(2, 4, 64, dtype=torch.bfloat16)
    k_nhd, v_nhd = torch.randn(3, 8, 2, 64, dtype=torch.bfloat16), torch.randn(3, 8, 2, 64, dtype=torch.bfloat16)
    ws = torch.empty(1 << 20, dtype=torch.uint8)
    table, lens = torch.tensor([[0, 1], [2, 0]], dtype=torch.int32), torch.tensor([12, 5], dtype=torch.int32)
    a = fi.decode.xqa_batch_decode_with_kv_cache(q, (k_nhd, v_nhd), ws, table, lens, 12, bmm1_scale=0.125)
    b = fi.decode.trtllm_batch_decode_with_kv_cache(q, (k_nhd.transpose(1, 2).contiguous(), v_nhd.transpose(1, 2).contiguous()), ws, table,
                                                    lens, 12, bmm1_scale=0.125)
    torch.testing.assert_close(a.float(), b.float(), atol=1e-2, rtol=1e-2)


def test_logits_processor_accepts_rules_and_checks_written_for_the_reference():
    """Reference-style extension points: FusionRule(pattern=<op classes>, guard=, build=, prio=), validity checks over the legalised op
    list, the reference's op class and builder names."""
    from flashinfer_b200.logits_processor import LogitsPipe, MinP, Sample, Softmax, Temperature, TopK, TopP
    from flashinfer_b200.logits_processor.fusion_rules import (FusionRule, build_topk_sampling, get_default_fusion_rules,
                                                              joint_topk_topp_sampleprobs_guard)
    from flashinfer_b200.logits_processor.legalization import validate_processor_chain
    from flashinfer_b200.logits_processor.operators import (FusedProbsTopKSampleOp, FusedProbsTopKTopPSampleOp, FusedTemperatureSoftmaxOp,
                                                           ProbsSampleOp, ProbsTopKOp, SoftmaxOp, TemperatureOp, TopPOp)
    from flashinfer_b200.logits_processor.types import LegalizationError, TensorType
    from flashinfer_b200.logits_processor.validators import (CompileError, get_default_validity_checks, indices_terminal_rule,
                                                            single_softmax_rule)

    assert len(get_default_fusion_rules()) == 6 and [c.__name__ for c in get_default_validity_checks()] == ["single_softmax_rule", "indices_terminal_rule"]
    seen = []
    rule = FusionRule(pattern=(TemperatureOp, SoftmaxOp), guard=lambda w: seen.append(len(w)) or False, build=lambda w: w[0], prio=1000)
    pipe = LogitsPipe([Temperature(), Softmax(), TopK(), Sample()], custom_fusion_rules=[rule])
    assert seen and isinstance(pipe.compiled_ops[0], FusedTemperatureSoftmaxOp) and isinstance(pipe.compiled_ops[1], FusedProbsTopKSampleOp)
    joint = LogitsPipe([TopK(joint_topk_topp=True), TopP(), Sample()], input_type=TensorType.PROBS)
    assert isinstance(joint.compiled_ops[0], FusedProbsTopKTopPSampleOp) and joint_topk_topp_sampleprobs_guard(joint.ops)
    assert isinstance(build_topk_sampling(LogitsPipe([TopK(), Sample()], input_type=TensorType.PROBS, compile=False).ops), FusedProbsTopKSampleOp)

    def no_top_p(ops):                                  # a check in the reference's form: it sees Op objects
        assert all(hasattr(o, "OUT") for o in ops)
        if any(isinstance(o, TopPOp) for o in ops):
            raise CompileError("top-p is not allowed here")

    LogitsPipe([Softmax(), TopK(), Sample()], custom_validity_checks=[no_top_p])
    with pytest.raises(CompileError):
        LogitsPipe([Softmax(), TopP(), Sample()], custom_validity_checks=[no_top_p])
    ops = LogitsPipe([Temperature(), Softmax(), TopK(), Sample()], compile=False).ops
    assert isinstance(ops[2], ProbsTopKOp) and isinstance(ops[3], ProbsSampleOp)
    single_softmax_rule(ops)
    indices_terminal_rule(ops)
    with pytest.raises(CompileError):
        single_softmax_rule(ops + [ops[1]])
    with pytest.raises(CompileError):
        indices_terminal_rule(ops + [ops[2]])
    validate_processor_chain([MinP(), Sample()])         # input type inferred from a probs-only processor
    with pytest.raises(LegalizationError):
        validate_processor_chain([])
    with pytest.raises(LegalizationError):
        validate_processor_chain([Sample(), Softmax()])


def test_logits_pipe_compile_time_hooks_of_the_reference():
    """compile(custom_fusion_rules, custom_validity_checks), Compiler.register_*, op-level validate_pipeline, TaggedTensor accessors."""
    import torch
    from flashinfer_b200.logits_processor import LogitsPipe, Sample, Softmax, TaggedTensor, Temperature, TensorType, TopK
    from flashinfer_b200.logits_processor.compiler import Compiler, compile_pipeline
    from flashinfer_b200.logits_processor.legalization import legalize_processors
    from flashinfer_b200.logits_processor.types import CompileError
    from flashinfer_b200.logits_processor.validators import validate_pipeline

    pipe = LogitsPipe([Temperature(), Softmax(), TopK(), Sample()], compile=False)
    seen = []
    pipe.compile(custom_validity_checks=[lambda ops: seen.append(len(ops))])
    assert seen and pipe.compiled_ops

    def no_topk(ops):
        if any("topk" in type(o).__name__.lower() for o in ops):
            raise CompileError("top-k is not allowed here")

    with pytest.raises(ValueError, match="Compilation failed: top-k"):
        pipe.compile(custom_validity_checks=[no_topk])
    ops = legalize_processors([Temperature(), Softmax(), TopK()])                       # initial_type defaults to logits
    assert ops == legalize_processors([Temperature(), Softmax(), TopK()], initial_type=TensorType.LOGITS) or len(ops) > 0
    validate_pipeline(ops)
    with pytest.raises(CompileError):
        validate_pipeline(ops, custom_checks=[no_topk])
    with pytest.raises(CompileError):
        validate_pipeline([])
    c = Compiler()
    c.register_validity_check(no_topk)
    with pytest.raises(CompileError):
        c.compile(ops)
    with pytest.raises(CompileError):
        Compiler().compile([])
    assert len(compile_pipeline(ops, None, None)) <= len(ops)
    t = TaggedTensor.indices(torch.zeros(2, 3, dtype=torch.int32))
    assert t.type == TensorType.INDICES and t.shape == (2, 3) and t.dtype == torch.int32 and t.size(1) == 3 and t.device.type == "cpu"
