"""@flashinfer_api instrumentation of the whole public surface (reference flashinfer/api_logging.py: levels 1/3/5, crash-safe
dumps, replay_from_dump / replay_sequence; tests/utils/test_logging.py, test_logging_replay.py)."""
import os
import subprocess
import sys

import pytest
import torch

import flashinfer_b200 as fi
from flashinfer_b200 import api_logging as al

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(code, **env):
    e = dict(os.environ, PYTHONPATH=ROOT, **{k: str(v) for k, v in env.items()})
    return subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True, timeout=300)


def test_every_public_op_is_registered():
    apis = set(al.registered_apis())
    for name in ("norm.rmsnorm", "norm.fused_add_rmsnorm", "rope.apply_rope", "page.append_paged_kv_cache", "sampling.top_k_sampling_from_probs",
                 "decode.single_decode_with_kv_cache", "prefill.single_prefill_with_kv_cache", "gemm.dense.mm_bf16", "gemm.lowp.mm_fp4",
                 "fused_moe.core.trtllm_fp4_block_scale_moe", "fused_moe.core.cutlass_fused_moe", "quantization.fp4.nvfp4_quantize",
                 "decode.BatchDecodeWithPagedKVCacheWrapper.plan", "decode.BatchDecodeWithPagedKVCacheWrapper.run",
                 "prefill.BatchPrefillWithPagedKVCacheWrapper.run", "mla._core.BatchMLAPagedAttentionWrapper.run",
                 "cascade.merge_state", "topk.top_k", "gemm.decode_linear.decode_linear"):
        assert name in apis, name
    assert len(apis) > 200
    # level 0 (default): the public names ARE the undecorated functions - zero overhead
    if al._LEVEL == 0:
        assert fi.rmsnorm is al._ORIGINALS["norm.rmsnorm"][2]


@pytest.mark.parametrize("level", [1, 3, 5])
def test_levels_via_environment(level):
    code = ("import torch, flashinfer_b200 as fi\n"
            "x = torch.randn(4, 64); w = torch.ones(64)\n"
            "fi.rmsnorm(x, w)\n"
            "fi.silu_and_mul(torch.randn(2, 8))\n")
    r = _run(code, FLASHINFER_LOGLEVEL=level, FLASHINFER_LOGDEST="stdout")
    assert r.returncode == 0, r.stderr
    assert "norm.rmsnorm" in r.stdout and "activation.silu_and_mul" in r.stdout
    if level >= 3:
        assert "shape=(4, 64)" in r.stdout and "-> Tensor(" in r.stdout
    else:
        assert "shape=" not in r.stdout
    assert ("min=" in r.stdout) == (level >= 5)


def test_wrapper_methods_are_logged_and_set_level_roundtrip(capsys):
    al.set_level(3, dest="stdout")
    try:
        q = torch.randn(8, 64)
        k = torch.randn(33, 2, 64)
        fi.single_decode_with_kv_cache(q, k, k.clone())
        out = capsys.readouterr().out
        assert "decode.single_decode_with_kv_cache(" in out
    finally:
        al.set_level(0)
    assert fi.single_decode_with_kv_cache is al._ORIGINALS["decode.single_decode_with_kv_cache"][2]


def test_dump_and_replay_cpu(tmp_path):
    code = ("import torch, flashinfer_b200 as fi\n"
            "torch.manual_seed(0)\n"
            "x = torch.randn(4, 64); r = torch.randn(4, 64); w = torch.rand(64) + 0.5\n"
            "fi.fused_add_rmsnorm(x, r, w, 1e-5)\n"
            "fi.merge_state(torch.randn(3, 2, 16), torch.randn(3, 2), torch.randn(3, 2, 16), torch.randn(3, 2))\n")
    r = _run(code, FLASHINFER_LOGLEVEL=3, FLASHINFER_DUMP_DIR=str(tmp_path), FLASHINFER_LOGDEST=str(tmp_path / "log.txt"))
    assert r.returncode == 0, r.stderr
    dumps = sorted(d for d in os.listdir(tmp_path) if os.path.isdir(tmp_path / d))
    assert len(dumps) == 2 and dumps[0].endswith("norm_fused_add_rmsnorm") and dumps[1].endswith("cascade_merge_state")
    res = al.replay_sequence(str(tmp_path), device="cpu")
    assert [x["api"] for x in res] == ["norm.fused_add_rmsnorm", "cascade.merge_state"]
    assert res[1]["match"] is True


@pytest.mark.gpu
def test_level5_stats_and_dump_replay_gpu(tmp_path):
    code = ("import torch, flashinfer_b200 as fi\n"
            "torch.manual_seed(0)\n"
            "x = torch.randn(64, 4096, device='cuda', dtype=torch.bfloat16); w = torch.ones(4096, device='cuda', dtype=torch.bfloat16)\n"
            "fi.rmsnorm(x, w)\n"
            "a = torch.randn(64, 512, device='cuda', dtype=torch.bfloat16); b = torch.randn(256, 512, device='cuda', dtype=torch.bfloat16)\n"
            "fi.mm_bf16(a, b.t())\n"
            "torch.cuda.synchronize()\n")
    r = _run(code, FLASHINFER_LOGLEVEL=5, FLASHINFER_DUMP_DIR=str(tmp_path), FLASHINFER_LOGDEST="stdout")
    assert r.returncode == 0, r.stderr
    assert "min=" in r.stdout and "nan=0" in r.stdout and "gemm.dense.mm_bf16" in r.stdout
    res = al.replay_sequence(str(tmp_path), device="cuda")
    assert all(x.get("match", True) for x in res) and len(res) >= 2
