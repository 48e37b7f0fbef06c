"""flashinfer_b200.cute_dsl: the reference's CuTe-DSL import path served by the native kernels (reference flashinfer/cute_dsl/,
tests/attention/test_cute_dsl_*.py).  CPU part: layout views, variant lowering, wrapper numerics through the eager paths."""
import math

import pytest
import torch

from flashinfer_b200 import cute_dsl, jit
from flashinfer_b200.cute_dsl.attention import (ALiBiAttention, AttentionWithSink, BatchMLADecodeCuteDSLWrapper, BatchPrefillCuteDSLWrapper,
                                                RPEAttention, SigmoidAttention, SoftCappingAttention, StandardAttention, cute_dsl_mla_decode,
                                                mla_get_split_kv, mla_get_workspace_size)


def test_mma_sf_layout_is_a_view_of_the_swizzled_buffer():
    from flashinfer_b200.quantization.fp4 import block_scale_interleave

    m, k, groups = 256, 320, 3
    kc = k // 16
    lin = (torch.arange(groups * m * kc) % 251).to(torch.uint8).view(groups, m, kc)
    swz = block_scale_interleave(lin).reshape(-1)
    v6 = cute_dsl.convert_sf_to_mma_layout(swz, m, k, groups)
    assert tuple(v6.shape) == cute_dsl.get_mma_sf_shape(m, k, groups) == (32, 4, 2, 4, 5, 3)
    assert v6.data_ptr() == swz.data_ptr() and not v6.is_contiguous()
    for g, r, c in [(0, 0, 0), (1, 37, 3), (2, 200, 19), (1, 255, 12)]:
        assert int(v6[r % 32, (r % 128) // 32, r // 128, c % 4, c // 4, g]) == int(lin[g, r, c])
    back = cute_dsl.convert_sf_from_mma_layout(v6, m, k, groups)
    assert tuple(back.shape) == (groups * 256, 20) and torch.equal(back.reshape(-1), swz)
    with pytest.raises(ValueError):
        cute_dsl.convert_sf_to_mma_layout(swz[:-1], m, k, groups)


def test_alibi_slope_schedule():
    from flashinfer_b200.utils import get_alibi_slopes

    assert torch.allclose(ALiBiAttention.get_slopes(8), get_alibi_slopes(8).cpu().float())
    s12 = ALiBiAttention.get_slopes(12)
    assert s12.numel() == 12 and torch.allclose(s12[:8], ALiBiAttention.get_slopes(8)) and torch.allclose(s12[8:], ALiBiAttention.get_slopes(16)[0::2][:4])


def _reference(q, k, v, ind, var, sm):
    out = torch.zeros(q.shape[0], q.shape[1], v.shape[-1])
    g = q.shape[1] // k.shape[1]
    for i in range(ind.numel() - 1):
        s, e = int(ind[i]), int(ind[i + 1])
        n = e - s
        kf, vf = k[s:e].float().repeat_interleave(g, 1), v[s:e].float().repeat_interleave(g, 1)
        lg = torch.einsum("qhd,khd->hqk", q[s:e].float(), kf) * sm
        pos = torch.arange(n)
        if isinstance(var, SoftCappingAttention):
            lg = var.cap * torch.tanh(lg / var.cap)
        if isinstance(var, ALiBiAttention):
            lg = lg + var.alibi_slopes[:, None, None] * (pos[None, None, :] - pos[None, :, None])
        if isinstance(var, SigmoidAttention):
            p = torch.sigmoid(lg * var.scale + var.bias).masked_fill(pos[None, :] > pos[:, None], 0.0)
        else:
            lg = lg.masked_fill(pos[None, :] > pos[:, None], float("-inf"))
            if isinstance(var, AttentionWithSink):
                p = torch.softmax(torch.cat([lg, var.sinks[:, None, None].expand(-1, n, 1)], -1), -1)[..., :-1]
            else:
                p = torch.softmax(lg, -1)
        out[s:e] = torch.einsum("hqk,khd->qhd", p, vf)
    return out


@pytest.mark.parametrize("make", [lambda: None, StandardAttention, lambda: ALiBiAttention(torch.tensor([0.3, 0.1, 0.05, 0.7])),
                                  lambda: SoftCappingAttention(5.0), lambda: AttentionWithSink(torch.tensor([0.5, -1.0, 2.0, 0.0])),
                                  lambda: SigmoidAttention(0.7, -0.2)],
                         ids=["none", "standard", "alibi", "softcap", "sink", "sigmoid"])
def test_prefill_wrapper_variants_cpu(make):
    torch.manual_seed(0)
    ind = torch.tensor([0, 5, 14], dtype=torch.int32)
    q, k, v = (torch.randn(14, h, 64).to(torch.bfloat16) for h in (4, 2, 2))
    var = make()
    w = BatchPrefillCuteDSLWrapper(torch.empty(1 << 20, dtype=torch.uint8))
    with pytest.raises(RuntimeError):
        w.run(q, k, v)
    w.plan(ind, ind, 4, 2, 64, causal=True, sm_scale=0.125, q_data_type=torch.bfloat16, kv_data_type=torch.bfloat16, variant=var)
    out = torch.empty(14, 4, 64, dtype=torch.bfloat16)
    assert w.run(q, k, v, out=out) is out
    torch.testing.assert_close(out.float(), _reference(q, k, v, ind, var, 0.125), atol=2e-2, rtol=2e-2)


def test_rpe_variant_lowers_to_a_compiled_hook():
    table = torch.randn(4, 2 * 8 + 1)
    var = RPEAttention(table, 8)
    assert var.is_compiled_hook and list(var.tensors) == ["rpe_table"] and list(var.scalars) == ["rpe_max_rel_dist"]
    args = var.jit_args("unit_rpe", torch.bfloat16, torch.bfloat16, 128, 128)
    assert args[7:11] == [["rpe_table"], ["float"], ["rpe_max_rel_dist"], ["double"]] and args[11] == "RPEAttention"
    assert var.run_args()[1] == 8.0 and var.run_args()[0].dtype == torch.float32
    with pytest.raises(ValueError):
        RPEAttention(torch.randn(4, 16), 8)
    if not jit.have_nvcc():
        pytest.skip("nvcc not available")
    spec = jit.gen_customize_batch_prefill_module("auto", *args)
    try:
        jit.build_module(spec)                         # the hook compiles into the tcgen05 prefill kernel for sm_100a
        assert spec.is_fresh()
    finally:
        for p in (spec.so_path, spec.hash_path):
            if p.exists():
                p.unlink()


def test_mla_wrapper_cpu_matches_reference():
    torch.manual_seed(1)
    b, h, page = 3, 4, 16
    lens = torch.tensor([20, 7, 33], dtype=torch.int32)
    per = [(int(n) + page - 1) // page for n in lens]
    table = torch.zeros(b, max(per), dtype=torch.int32)
    ids = torch.randperm(sum(per) + 1)[: sum(per)].int()
    o = 0
    for i, p in enumerate(per):
        table[i, :p] = ids[o:o + p]
        o += p
    kv = torch.randn(sum(per) + 1, page, 576).clamp(-1, 1).to(torch.bfloat16)
    q = (torch.randn(b, 1, h, 576) * 0.5).to(torch.bfloat16)
    sm = 1.0 / math.sqrt(192.0)
    ws = torch.zeros(1 << 20, dtype=torch.int8)
    w = BatchMLADecodeCuteDSLWrapper(ws)
    with pytest.raises(RuntimeError):
        w.run(q, kv, table, lens, int(lens.max()), sm)
    w.plan(512, 64, h, page, torch.bfloat16)
    got = w.run(q, kv, table, lens, int(lens.max()), sm, output_scale=0.5)
    got_fn = cute_dsl_mla_decode(q, kv, ws, 512, 64, table, lens, int(lens.max()), sm, 0.5)
    ref = torch.zeros(b, 1, h, 512)
    for i in range(b):
        n = int(lens[i])
        rows = kv[table[i, : per[i]].long()].reshape(-1, 576)[:n].float()
        ref[i, 0] = torch.softmax(q[i, 0].float() @ rows.t() * sm, -1) @ rows[:, :512] * 0.5
    torch.testing.assert_close(got.float(), ref, atol=2e-2, rtol=2e-2)
    torch.testing.assert_close(got_fn.float(), got.float(), atol=0, rtol=0)
    with pytest.raises(NotImplementedError):
        w.plan(512, 64, h, page, torch.bfloat16, variant=SoftCappingAttention(30.0))
    with pytest.raises(TypeError):
        BatchMLADecodeCuteDSLWrapper(torch.zeros(16, dtype=torch.float32))


def test_mla_split_heuristics():
    assert mla_get_split_kv(128, 1, 8192, 148) == 1                     # enough requests to fill the SM pairs
    s = mla_get_split_kv(4, 1, 8192, 148)
    assert 1 < s <= 74 // 4
    assert mla_get_split_kv(1, 1, 64, 148) == 1                          # one page: nothing to split
    assert mla_get_workspace_size(4, 1, 128, 512, 1) == 0
    assert mla_get_workspace_size(4, 1, 128, 512, 8) == 4 * 128 * 8 * 513 * 4


def test_package_names():
    for n in cute_dsl.__all__:
        assert hasattr(cute_dsl, n), n
    assert cute_dsl.is_cute_dsl_available() and cute_dsl.get_sm_version() >= 100
    ref, sf6 = cute_dsl.create_scale_factor_tensor(2, 128, 64, 16, device="cpu")
    assert tuple(ref.shape) == (128, 64, 2) and tuple(sf6.shape) == (32, 4, 1, 4, 1, 2)
    assert float(ref[5, 17, 1]) == float(sf6[5, 0, 0, 1, 0, 1].view(torch.float8_e4m3fn).float())
