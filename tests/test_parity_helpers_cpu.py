"""CPU checks of the host-side / reference paths of ops added late in the round (their CUDA kernels have GPU tests)."""
import math

import pytest
import torch

import flashinfer_b200 as fi


def test_fp8_group_quantize_cpu_roundtrip():
    from flashinfer_b200.gemm.lowp import fp8_group_quantize

    torch.manual_seed(0)
    x = torch.randn(6, 256) * 3
    q, sc = fp8_group_quantize(x)
    assert q.dtype == torch.float8_e4m3fn and sc.shape == (6, 2)
    dq = (q.float().view(6, 2, 128) * sc[..., None]).view(6, 256)
    assert (dq - x).abs().max().item() < 0.07 * x.abs().max().item()
    # gated: rows are [linear | gate]
    h = torch.randn(4, 512)
    qg, sg = fp8_group_quantize(h, gated=True)
    ref = h[:, :256] * torch.nn.functional.silu(h[:, 256:])
    dqg = (qg.float().view(4, 2, 128) * sg[..., None]).view(4, 256)
    assert (dqg - ref).abs().max().item() < 0.07 * ref.abs().max().item() + 1e-3


def test_linear_gated_silu_cpu_matches_unfused():
    from flashinfer_b200.gemm import interleave_gate_up, linear_gated_silu

    x, w = torch.randn(7, 64), torch.randn(48, 64)
    ref = fi.silu_and_mul(x @ w.t())
    torch.testing.assert_close(linear_gated_silu(x, interleave_gate_up(w)), ref, rtol=1e-5, atol=1e-5)


def test_fp8_block_moe_cpu_fallback_matches_dequant():
    from flashinfer_b200.fused_moe import moe_forward, moe_forward_fp8_block, route
    from flashinfer_b200.fused_moe.core import _dequant_fp8_block

    torch.manual_seed(1)
    T, E, K, H, I = 16, 4, 2, 128, 128
    x = torch.randn(T, H).bfloat16()

    def q(wt):
        Eb, N, Kd = wt.shape
        blk = wt.float().reshape(Eb, N // 128, 128, Kd // 128, 128)
        s = blk.abs().amax((2, 4)) / 448.0
        return (blk / s[:, :, None, :, None]).reshape(Eb, N, Kd).to(torch.float8_e4m3fn), s

    w1q, s1 = q(torch.randn(E, 2 * I, H) / H ** 0.5)
    w2q, s2 = q(torch.randn(E, H, I) / I ** 0.5)
    ids, w = route(torch.randn(T, E), None, K, 1)
    out = moe_forward_fp8_block(x, None, ids, w, w1q, s1, w2q, s2)
    ref = moe_forward(x, ids, w, _dequant_fp8_block(w1q, s1), _dequant_fp8_block(w2q, s2))
    torch.testing.assert_close(out.float(), ref.float(), rtol=2e-2, atol=2e-2)


def test_deepgemm_wrappers_cpu():
    from flashinfer_b200.gemm import batch_deepgemm_fp8_nt_groupwise, group_deepgemm_fp8_nt_groupwise

    torch.manual_seed(2)
    E, N, K = 2, 128, 256
    m_idx = torch.cat([torch.zeros(128), torch.ones(256)]).int()
    a = torch.randn(384, K)
    g = a.view(384, K // 128, 128)
    a_s = g.abs().amax(-1) / 448.0
    aq = (g / a_s[..., None]).view(384, K).to(torch.float8_e4m3fn)
    w = torch.randn(E, N, K) / K ** 0.5
    blk = w.view(E, 1, 128, K // 128, 128)
    w_s = blk.abs().amax((2, 4)) / 448.0
    wq = (blk / w_s[:, :, None, :, None]).reshape(E, N, K).to(torch.float8_e4m3fn)
    a_dq = (aq.float().view(384, K // 128, 128) * a_s[..., None]).view(384, K)
    w_dq = (wq.float().view(E, 1, 128, K // 128, 128) * w_s[:, :, None, :, None]).reshape(E, N, K)
    ref = torch.cat([a_dq[:128] @ w_dq[0].t(), a_dq[128:] @ w_dq[1].t()])
    out = group_deepgemm_fp8_nt_groupwise(aq, wq, a_s, w_s, m_idx)
    torch.testing.assert_close(out.float(), ref, rtol=3e-2, atol=3e-2)
    masked = torch.tensor([100, 128], dtype=torch.int32)
    o = batch_deepgemm_fp8_nt_groupwise(aq[:256].view(2, 128, K), wq, a_s[:256].view(2, 128, -1), w_s, masked)
    torch.testing.assert_close(o[0, :100].float(), (a_dq[:100] @ w_dq[0].t()), rtol=3e-2, atol=3e-2)


def test_topk_cluster_entry_points_cpu():
    from flashinfer_b200 import topk

    x = torch.randn(3, 100)
    idx = topk.topk_clusters_exact(x, 5)
    assert idx.dtype == torch.int32 and idx.shape == (3, 5)
    assert set(idx[0].tolist()) == set(torch.topk(x[0], 5).indices.tolist())
    idx2, vals = topk.topk_clusters_exact(x, 5, output_values=True, out_dtype=torch.int64)
    assert idx2.dtype == torch.int64 and torch.allclose(vals.sort(-1).values, torch.topk(x, 5).values.sort(-1).values)
    assert topk.get_fast_topk_clusters(8) == 8 and topk.roundup_kbyte(1) == 1024 and topk.can_implement_filtered_topk()


def test_misc_parity_helpers_cpu():
    from flashinfer_b200 import norm, sparse
    from flashinfer_b200.fused_moe import MoEInputs, convert_to_block_layout

    m = torch.rand(5, 2, 3) > 0.5
    flat = sparse.convert_bsr_mask_layout(m, torch.tensor([0, 2, 5]))
    assert flat.shape == (30,) and torch.equal(flat[:12], m[:2].transpose(0, 1).reshape(-1))
    q, k = torch.randn(4, 2, 8), torch.randn(4, 1, 8)
    norm.qk_rmsnorm_cute(q, k, torch.ones(8), torch.ones(8))
    assert (q.pow(2).mean(-1) - 1).abs().max().item() < 1e-3 and (k.pow(2).mean(-1) - 1).abs().max().item() < 1e-3
    w = torch.arange(2 * 4 * 8).float().view(2, 4, 8)
    b = convert_to_block_layout(w, 4)
    assert b.shape == (2, 2, 4, 4) and torch.equal(b[0, 1, 2], w[0, 2, 4:8])
    mi = MoEInputs(hidden_states=torch.zeros(1))
    assert MoEInputs.from_list(mi.to_list()).hidden_states is mi.hidden_states and MoEInputs.idx("output") == 0


def test_fmha_v2_prefill_deepseek_cpu():
    from flashinfer_b200 import prefill, reference

    torch.manual_seed(0)
    q, k, v = torch.randn(2, 33, 2, 192), torch.randn(2, 33, 2, 192), torch.randn(2, 33, 2, 128)
    out = torch.empty(2, 33, 2, 128)
    o, lse = prefill.fmha_v2_prefill_deepseek(q, k, v, out, 2, 192, 33, 0.0, return_lse=True)
    for b in range(2):
        ref, _ = reference.attention_ref(q[b], k[b], v[b], True)
        torch.testing.assert_close(o[b], ref, rtol=1e-4, atol=1e-4)
    assert lse.shape == (2, 33, 2)


def test_user_jit_spec_builds_and_loads(tmp_path):
    """gen_jit_spec: users compile their own .cu with the package toolchain (sm_100a flags, fib200 headers, uniform C ABI)."""
    import pytest

    from flashinfer_b200 import jit

    if not jit.have_nvcc():
        pytest.skip("nvcc not available")
    src = tmp_path / "user_axpy.cu"
    src.write_text(
        '#include <fib200/common.cuh>\n'
        'using namespace fib200;\n'
        'FIB_EXPORT_LAST_ERROR()\n'
        '__global__ void axpy_kernel(const float* x, float* y, float a, int64_t n) {\n'
        '  int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;\n'
        '  if (i < n) y[i] += a * x[i];\n'
        '}\n'
        'extern "C" int user_axpy(void* x, void* y, double a, int64_t n, int64_t stream) {\n'
        '  if (n == 0) return 0;\n'
        '  axpy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>((const float*)x, (float*)y, (float)a, n);\n'
        '  FIB_CUDA_CHECK(cudaGetLastError());\n'
        '  return 0;\n'
        '}\n')
    spec = jit.gen_jit_spec("user_axpy_test", [src])
    try:
        assert spec.status == jit.JitSpecStatus.NOT_COMPILED
        mod = spec.build_and_load()
        assert spec.status == jit.JitSpecStatus.COMPILED and mod.has("user_axpy")
        assert jit.jit_spec_registry.get_spec_status("user_axpy_test").status == jit.JitSpecStatus.COMPILED
        mod.call("user_axpy", None, None, 2.0, 0, 0)  # n = 0: returns before touching the GPU
    finally:
        jit.clear_cache_dir()
        jit.REGISTRY.pop("user_axpy_test", None)
    assert not spec.so_path.exists()


def test_utils_helper_names_cpu():
    import pytest

    from flashinfer_b200 import utils as u

    assert torch.equal(u.get_indptr(torch.tensor([2, 0, 3])), torch.tensor([0, 2, 2, 5]))
    s = u.get_alibi_slopes(12)
    assert s.shape == (12,) and abs(float(s[0]) - 2 ** -1) < 1e-6 and abs(float(s[8]) - 2 ** -0.5) < 1e-6
    assert u.canonicalize_torch_dtype("bfloat16") is torch.bfloat16 and u.is_float8(torch.zeros(1).to(torch.float8_e4m3fn))
    assert u.calculate_tile_tokens_dim(1024, 256, 8) == 64 and u.calculate_tile_tokens_dim(1, 256, 8) == 8
    with pytest.raises(ValueError):
        u.check_shape_dtype_device(torch.zeros(2, 3), (3, 2), None, None, "x")
    u.check_shape_dtype_device(torch.zeros(2, 3), (2, 3), torch.float32, torch.device("cpu"), "x")
    t = u.FP4Tensor(torch.zeros(4, 8, dtype=torch.uint8), torch.zeros(4, 1))
    assert t.original_shape == (4, 16)
    assert u.get_shuffle_block_size(128) == 32 and u.get_shuffle_block_size(64) == 16
    idx = u.get_shuffle_matrix_a_row_indices(torch.zeros(256, 4), 128)
    assert sorted(idx.tolist()) == list(range(256))
    assert u.version_at_least("2.11.0", "2.8") and u.determine_attention_backend(None, 0, False, False, None, None) == "sm100"

    @u.supported_compute_capability([100, 103])
    def f():
        return 1

    assert f.is_compute_capability_supported(100) and not f.is_compute_capability_supported(90)
    from flashinfer_b200 import tllm_enums as te

    assert int(te.DtypeTrtllmGen.Bfloat16) == (1 << 20) | (16 << 8) and te.trtllm_gen_dtype_has_scale(te.DtypeTrtllmGen.MxE2m1)
    assert te.deduce_trtllm_gen_tensor_dtype(torch.zeros(2, 8, dtype=torch.uint8), torch.zeros(2)) == te.DtypeTrtllmGen.E2m1
    from flashinfer_b200 import cuda_utils

    assert cuda_utils.checkCudaErrors((0, 5)) == 5 and cuda_utils.checkCudaErrors((0,)) is None
    with pytest.raises(RuntimeError):
        cuda_utils.checkCudaErrors((1, None))


def test_fused_moe_utils_buckets_and_swizzle():
    from flashinfer_b200.fused_moe import utils as u

    assert u.get_hybrid_num_tokens_buckets(10000) == (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 768, 1024, 1280, 1536, 1792, 2048,
                                                      2560, 3072, 3584, 4096, 8192, 10000)
    assert u.get_hybrid_num_tokens_buckets(300) == (1, 2, 4, 8, 16, 32, 64, 128, 256, 300)
    grid = u.get_hybrid_num_tokens_buckets(10000)
    for x in (1, 3, 255, 256, 257, 1000, 2048, 2049, 4000, 4097, 9999, 20000):
        b = u.map_to_hybrid_bucket(x, 10000)
        assert b in grid and (b >= x or b == 10000)
        assert b == min(g for g in grid if g >= min(x, 10000))       # the smallest grid point that covers x
    assert u.map_to_hybrid_bucket_uncapped(5000) == 8192 and u.map_to_hybrid_bucket(0, 64) == 1
    assert u.round_to_nearest_bucket(350, [100, 200, 500, 1000]) == 200
    assert u.round_to_nearest_bucket(350, [100, 200, 500, 1000], round_map=True) == 500
    assert u.round_to_nearest_bucket(2000, [100, 200, 500, 1000], round_map=True) == 1000
    assert u.make_bucket_mapper((500, 100, 100, 200))(5) == 100
    assert (u.next_positive_power_of_2(5), u.last_positive_power_of_2(5), u.last_positive_power_of_2(8)) == (8, 4, 8)
    sf = torch.randint(0, 255, (2, 200, 6), dtype=torch.uint8)
    sw = u.swizzle_sf(sf, 200, 96)
    assert sw.numel() == 2 * 256 * 8 and u.compute_swizzled_sf_shape(200, 6) == (256, 8)
    assert torch.equal(u.unswizzle_sf(sw, 200, 96).view(2, 200, 6), sf)
    sf2 = torch.randint(0, 255, (2, 128, 8), dtype=torch.uint8)
    parts = torch.cat([u.swizzle_sf(sf2[i], 128, 128) for i in range(2)])
    assert torch.equal(u.reswizzle_sf(parts, 128, 128), u.swizzle_sf(sf2.reshape(256, 8), 256, 128))
    assert u.get_fp4_shape([4, 100, 64], 16) == ([4, 100, 32], 2048)
    with u.model_extra_attrs({"a": 1}):
        assert u.get_model_extra_attrs() == {"a": 1}
    assert u.get_model_extra_attrs() is None


def test_cli_export_compile_commands_and_cubin_commands(tmp_path, capsys):
    import json

    from flashinfer_b200.__main__ import main

    out = tmp_path / "cc.json"
    assert main(["export-compile-commands", "norm", "pod_sm100", "-o", str(out)]) == 0
    entries = json.loads(out.read_text())
    assert len(entries) == 2 and all(e["file"].endswith(".cu") and "arch=compute_100a,code=sm_100a" in e["arguments"] for e in entries)
    assert entries[0]["arguments"][1:3] == ["-c", entries[0]["file"]]
    assert main(["list-cubins"]) == 0 and main(["download-cubin"]) == 0
    text = capsys.readouterr().out
    assert "gemm_sm100" in text and "built from source" in text


def test_module_accessors_and_moe_layout_helpers():
    import torch

    from flashinfer_b200 import decode, fused_moe, gemm, jit, mla, norm, prefill

    for get, name in ((decode.get_batch_decode_module, "decode_sm100"), (prefill.get_batch_prefill_module, "prefill_sm100"),
                      (norm.get_norm_module, "norm"), (gemm.get_gemm_sm100_module, "gemm_sm100"), (mla.get_batch_mla_module, "mla_sm100")):
        assert get("ignored", dtype=torch.bfloat16) is jit.load(name)
    assert fused_moe.get_reorder_rows_for_gated_act_gemm_row_indices(torch.zeros(8, 2)).tolist() == [0, 4, 1, 5, 2, 6, 3, 7]
    idx = fused_moe.get_w2_permute_indices_with_cache({}, torch.zeros(256, 16), 128)
    assert sorted(idx.tolist()) == list(range(256))
    calls = []
    cached = prefill.make_hashable_cache(lambda names, opts=None: calls.append(1) or len(calls))
    assert cached(["a", "b"], opts={"k": [1]}) == cached(["a", "b"], opts={"k": [1]}) == 1 and cached(["a"]) == 2


def test_moe_prepared_cache_is_tied_to_the_live_weight_tensors():
    """Load-time weight preparation is cached per weight tensor; an entry must die with the tensor it was computed from
    (the allocator hands the same address to the next tensor of that shape) and must survive fresh views of a live base."""
    import gc

    from flashinfer_b200.fused_moe import core

    calls = []

    def prep(t):
        return core._prepared("unit-test", [t], lambda: (calls.append(1), t.clone())[1])

    base = torch.nn.Parameter(torch.arange(64, dtype=torch.float32))
    first = prep(base.view(8, 8))                      # temporary view: the storage stays alive
    assert prep(base.view(8, 8)) is first and prep(base.data.view(8, 8)) is first and prep(base.detach().view(8, 8)) is first and len(calls) == 1
    base = base.data
    base.add_(1)                                       # in-place update of the weights -> new version -> recomputed
    assert prep(base.view(8, 8)) is not first and len(calls) == 2

    a = torch.full((32,), 1.0)
    ptr = a.data_ptr()
    got_a = prep(a)
    del a
    gc.collect()
    b = torch.full((32,), 2.0)
    if b.data_ptr() == ptr:                            # address reuse is what usually happens; the rule is checked when it does
        assert torch.equal(prep(b), b) and prep(b) is not got_a


def test_artifacts_describe_the_in_tree_libraries():
    """The artifact vocabulary of the reference mapped onto the compiled modules: checksums are the content hashes of the build."""
    import os

    from flashinfer_b200 import artifacts, jit

    status = dict(artifacts.get_artifacts_status())
    assert len(status) == len(jit.REGISTRY) and set(artifacts.get_checksums()) == set(status)
    name, digest = next(artifacts.get_subdir_file_list())
    assert name.endswith(".so") and len(digest) == 64
    assert set(artifacts.get_checksums(["comm"])) == {k for k in status if k.startswith("comm/")} != set()
    built = set(artifacts.get_available_cubin_files())
    assert built <= set(status) and all(status[k] for k in built if artifacts.CheckSumHash.recorded().get(k.split("/")[1][:-3]) ==
                                        artifacts.CheckSumHash.expected()[k.split("/")[1][:-3]])
    assert any(h.endswith(".cuh") for h in artifacts.get_available_header_files())
    assert artifacts.ArtifactPath().TRTLLM_GEN_FMHA == "" and artifacts.ArtifactPath.GEMM == "gemm"
    with artifacts.temp_env_var("FIB200_UNIT_ENV", "x"):
        assert os.environ["FIB200_UNIT_ENV"] == "x"
    assert "FIB200_UNIT_ENV" not in os.environ


def test_triton_path_helpers_on_native_ops():
    """flashinfer.triton.{activation,norm,cascade}: unscaled forms equal the native ops, scales follow the Triton kernels' semantics
    (inputs multiplied by in-scale, result multiplied by out-scale and clamped to the output dtype)."""
    from flashinfer_b200.triton import activation as ta
    from flashinfer_b200.triton import cascade as tc
    from flashinfer_b200.triton import norm as tn

    g = torch.Generator().manual_seed(0)
    x = torch.randn(5, 64, generator=g).to(torch.bfloat16)
    ref = torch.nn.functional.silu(x[:, :32].float()) * x[:, 32:].float()
    torch.testing.assert_close(ta.silu_and_mul(x).float(), ref, atol=2e-2, rtol=2e-2)
    xs, os_ = torch.tensor(0.5), torch.tensor(300.0)
    q = ta.silu_and_mul(x, xs, os_, torch.float8_e4m3fn)
    want = (torch.nn.functional.silu(x[:, :32].float() * 0.5) * (x[:, 32:].float() * 0.5) * 300.0).clamp(-448, 448).to(torch.float8_e4m3fn)
    assert q.dtype == torch.float8_e4m3fn and torch.equal(q.float(), want.float())
    with pytest.raises(TypeError):
        ta.scale_and_clamp(x, 1.0, torch.int8)

    w = (1 + 0.1 * torch.randn(64, generator=g)).to(torch.bfloat16)
    out = torch.empty_like(x)
    tn.rms_norm(x, w, out, 1e-6)
    nref = w.float() * x.float() * torch.rsqrt((x.float() ** 2).mean(-1, keepdim=True) + 1e-6)
    torch.testing.assert_close(out.float(), nref, atol=2e-2, rtol=2e-2)
    o8 = torch.empty(5, 64, dtype=torch.float8_e4m3fn)
    tn.rms_norm(x, w, o8, 1e-6, in_scale=torch.tensor(2.0), out_scale=torch.tensor(100.0))
    x2 = x.float() * 2
    want = (w.float() * x2 * torch.rsqrt((x2 ** 2).mean(-1, keepdim=True) + 1e-6) * 100).clamp(-448, 448).to(torch.float8_e4m3fn)
    assert torch.equal(o8.float(), want.float())
    xr, res = x.clone(), torch.randn(5, 64, generator=g).to(torch.bfloat16)
    r0 = res.clone()
    tn.rms_norm_add_residual(xr, res, w, 1e-6)
    summed = (x.float() + r0.float()).to(torch.bfloat16)
    torch.testing.assert_close(res.float(), summed.float(), atol=1e-2, rtol=1e-2)
    torch.testing.assert_close(xr.float(), w.float() * summed.float() * torch.rsqrt((summed.float() ** 2).mean(-1, keepdim=True) + 1e-6), atol=3e-2, rtol=3e-2)
    xo = torch.empty(5, 64, dtype=torch.float8_e4m3fn)
    res2 = r0.clone()
    tn.rms_norm_add_residual(x, res2, w, 1e-6, x_out=xo, x_in_scale=torch.tensor(0.5), x_out_scale=torch.tensor(50.0))
    assert torch.equal(res2, (r0.float() + 0.5 * x.float()).to(torch.bfloat16)) and xo.dtype == torch.float8_e4m3fn

    # segmented merge against merging each segment by hand
    v = torch.randn(9, 3, 8, generator=g)
    s = torch.randn(9, 3, generator=g) * 3
    indptr = torch.tensor([0, 4, 4, 5, 9], dtype=torch.int32)               # segment 1 is empty
    vo, so = tc.variable_length_merge_states(v, s, indptr)
    for i, (a, b) in enumerate(zip(indptr.tolist()[:-1], indptr.tolist()[1:])):
        if a == b:
            assert float(vo[i].abs().sum()) == 0.0 and bool(torch.isinf(so[i]).all())
            continue
        wts = torch.exp2(s[a:b] - torch.logsumexp(s[a:b] * math.log(2.0), 0) / math.log(2.0))
        torch.testing.assert_close(vo[i], (wts[..., None] * v[a:b]).sum(0), atol=1e-5, rtol=1e-5)
        torch.testing.assert_close(so[i], torch.logsumexp(s[a:b] * math.log(2.0), 0) / math.log(2.0), atol=1e-5, rtol=1e-5)
    v2, s2 = tc.merge_states(v[:4].unsqueeze(0).transpose(0, 1).reshape(1, 4, 3, 8), s[:4].reshape(1, 4, 3))
    torch.testing.assert_close(v2[0], vo[0], atol=1e-4, rtol=1e-4)


def test_dit_fusions_match_their_definitions():
    """diffusion_ops: gated residual + LayerNorm + modulation (reference norm/__init__.py:1057-1440) against plain PyTorch."""
    import flashinfer_b200.diffusion_ops as dit
    from flashinfer_b200 import norm

    assert norm.fused_dit_residual_layernorm_scale_shift is dit.fused_dit_residual_layernorm_scale_shift
    g = torch.Generator().manual_seed(0)
    b, s, h = 2, 5, 64
    r = lambda *sh: torch.randn(*sh, generator=g)  # noqa: E731
    x, res, gate = r(b, s, h).to(torch.bfloat16), r(b, s, h).to(torch.bfloat16), r(b, 1, h).to(torch.bfloat16)
    gamma, beta, scale, shift = r(h), r(h), r(b, 1, h) * 0.1, r(b, 1, h) * 0.1
    ln = lambda t: torch.nn.functional.layer_norm(t.float(), (h,), eps=1e-6)  # noqa: E731

    ro, no = dit.fused_dit_gate_residual_layernorm_gamma_beta(x, res, gate, gamma, beta, gate_bias=torch.ones(h))
    want_r = (res.float() + x.float() * (gate.float() + 1.0)).to(torch.bfloat16)
    torch.testing.assert_close(ro.float(), want_r.float(), atol=1e-2, rtol=1e-2)
    torch.testing.assert_close(no.float(), ln(want_r) * gamma + beta, atol=4e-2, rtol=4e-2)

    ro, no = dit.fused_dit_gate_residual_layernorm_scale_shift(x, res, gate, scale, shift, scale_bias=torch.zeros(h), shift_bias=torch.ones(h))
    want_r = (res.float() + x.float() * gate.float()).to(torch.bfloat16)
    torch.testing.assert_close(no.float(), ln(want_r) * (1 + scale) + shift + 1.0, atol=4e-2, rtol=4e-2)

    rbuf, nbuf = torch.empty_like(x), torch.empty_like(x)
    ro, no = dit.fused_dit_residual_layernorm_scale_shift(x, res, scale, shift, residual_out=rbuf, norm_out=nbuf, input_global_scaling_factor=0.5)
    assert ro is rbuf and no is nbuf
    want_r = (res.float() + 0.5 * x.float()).to(torch.bfloat16)
    torch.testing.assert_close(rbuf.float(), want_r.float(), atol=1e-2, rtol=1e-2)
    torch.testing.assert_close(nbuf.float(), ln(want_r) * (1 + scale) + shift, atol=4e-2, rtol=4e-2)

    _, q8, sf8 = dit.fused_dit_residual_layernorm_scale_shift(x, res, scale, shift, use_mxfp8=True)
    assert q8.dtype == torch.float8_e4m3fn and q8.shape == x.shape and sf8.dtype == torch.uint8
    _, q4, sf4 = dit.fused_dit_residual_layernorm_scale_shift(x, res, scale, shift, use_nvfp4=True, global_scaling_factor=torch.tensor([100.0]))
    assert q4.dtype == torch.uint8 and q4.shape == (b, s, h // 2)
