"""Autotuner host logic on CPU: tactic choice, bucket mapping + overrides, config file round trip (reference
tests/autotuner/test_autotuner_core.py, test_autotuner_tile_mismatch.py for the strategy)."""
import json
import time

import pytest
import torch

from flashinfer_b200.autotuner import (AutoTuner, ConstraintSpec, DynamicTensorSpec, TunableRunner, TuningConfig, autotune,
                                       is_in_profile_measurement)


class SleepRunner(TunableRunner):
    """Tactic t sleeps cost[t] ms; tactic 3 always fails."""
    cost = {-1: 3.0, 0: 2.0, 1: 0.2, 2: 1.0}

    def __init__(self):
        self.calls = []
        self.seen_measure_flag = []

    def get_valid_tactics(self, inputs, profile):
        return [0, 1, 2, 3]

    def forward(self, inputs, tactic=-1, do_preparation=False, **kwargs):
        self.calls.append((tuple(inputs[0].shape), tactic))
        self.seen_measure_flag.append(is_in_profile_measurement())
        if tactic == 3:
            raise RuntimeError("unsupported tile")
        time.sleep(self.cost[tactic] * 1e-3)
        return inputs[0] * 2


@pytest.fixture()
def tuner():
    t = AutoTuner.get()
    t.clear_cache()
    t.reset_statistics()
    yield t
    t.clear_cache()


def _cfg(buckets=()):
    return TuningConfig(dynamic_tensor_specs=(DynamicTensorSpec((0, 1), (0, 0), tuple(buckets)),),
                        constraint_specs=(ConstraintSpec(1, 1, lambda shapes: shapes[0][1] * 2),), use_cold_l2_cache=False)


def test_choose_one_profiles_only_inside_autotune(tuner):
    r = SleepRunner()
    x = [torch.zeros(5, 8), torch.zeros(5, 16)]
    runner, tac = tuner.choose_one("op", [r], _cfg(), x)
    assert tac == -1 and not r.calls and tuner.stats["misses"] == 1
    with autotune():
        runner, tac = tuner.choose_one("op", [r], _cfg(), x)
    assert runner is r and tac == 1 and tuner.stats["failed"] == 1 and tuner.stats["profiled"] == 3
    assert all(r.seen_measure_flag) and not is_in_profile_measurement()
    n = len(r.calls)
    # 5 and 7 share the bucket 8: served from the cache, outside tuning mode too
    runner, tac = tuner.choose_one("op", [r], _cfg(), [torch.zeros(7, 8), torch.zeros(7, 16)])
    assert tac == 1 and len(r.calls) == n and tuner.stats["hits"] == 1
    # a different static dim is a different key
    assert tuner.choose_one("op", [r], _cfg(), [torch.zeros(7, 4), torch.zeros(7, 8)])[1] == -1


def test_gen_tuning_buckets_are_all_profiled(tuner):
    r = SleepRunner()
    with autotune():
        tuner.choose_one("op", [r], _cfg((2, 64)), [torch.zeros(5, 8), torch.zeros(5, 16)])
    keys = {k[2] for k in tuner.profiling_cache}
    assert keys == {((8, 8), (8, 16)), ((2, 8), (2, 16)), ((64, 8), (64, 16))}
    assert {c[0][0] for c in r.calls} == {5, 2, 64}            # live shape + one synthetic input per listed bucket


def test_bucket_overrides_nest_and_round(tuner):
    spec = DynamicTensorSpec((0,), (0,), (16, 32, 64))
    assert tuner.get_effective_map_to_tuning_buckets(spec)(200) == 256
    with autotune(False, tuning_buckets=(512, 128, 128, 256)):
        m = tuner.get_effective_map_to_tuning_buckets(spec)
        assert (m(200), m(100), m(5000)) == (128, 128, 512)     # floor, clamped
        with autotune(False, round_up=True):
            m = tuner.get_effective_map_to_tuning_buckets(spec)
            assert (m(200), m(100), m(5000)) == (256, 128, 512)  # buckets inherited, ceil
        assert tuner.get_effective_map_to_tuning_buckets(spec)(200) == 128
    with autotune(False, round_up=True):
        assert tuner.get_effective_map_to_tuning_buckets(spec)(20) == 32   # ceil over the spec's own buckets
    assert tuner.get_effective_map_to_tuning_buckets(spec)(200) == 256
    with pytest.raises(ValueError):
        with autotune(tuning_buckets=()):
            pass


def test_config_file_roundtrip_and_merge(tuner, tmp_path):
    path = str(tmp_path / "tuned.json")
    r = SleepRunner()
    with autotune(cache=path):
        tuner.choose_one("op_a", [r], _cfg(), [torch.zeros(5, 8), torch.zeros(5, 16)], extras=("bf16", (1, 2)))
    data = json.load(open(path))
    assert data["metadata"]["device"] == "cpu" and len(data["configs"]) == 1
    tuner.clear_cache()
    with autotune(cache=path):                                   # second process: adds op_b, must keep op_a
        tuner.clear_cache()
        tuner.choose_one("op_b", [r], _cfg(), [torch.zeros(9, 8), torch.zeros(9, 16)])
    fresh = AutoTuner()
    assert fresh.load_configs(path) == 2
    hit, rid, tac = fresh.search_cache("op_a", [r], ((8, 8), (8, 16)), ("bf16", (1, 2)))
    assert hit and tac == 1
    assert fresh.search_cache("op_b", [r], ((16, 8), (16, 16)))[0]
    # a file from another device loads with a warning
    data["metadata"]["device"] = "NVIDIA H100"
    json.dump(data, open(path, "w"))
    with pytest.warns(RuntimeWarning):
        AutoTuner().load_configs(path)


def test_lowp_gemm_asks_the_tuner_for_its_n_tile(tuner, monkeypatch):
    """The block-scaled GEMM wrapper consults the tuner only when it is tuning or holds configs, passes the chosen N tile
    to the launcher and keys the choice by the bucketed M (the native launch is replaced: no GPU here)."""
    from flashinfer_b200.gemm import lowp

    seen = []

    def fake_raw(kind, a, b_nk, out, sfa, sfb, alpha_a, alpha_b, K, bn, tile_expert, meta, row_map):
        seen.append(bn)
        time.sleep({0: 2e-3, 64: 3e-3, 128: 2e-4, 192: 1e-3, 256: 1e-3}[bn])
        return out

    monkeypatch.setattr(lowp, "_launch_raw", fake_raw)
    a, b, o = torch.zeros(1, 48, 64, dtype=torch.uint8), torch.zeros(1, 512, 64, dtype=torch.uint8), torch.zeros(1, 48, 512)
    lowp._launch("fp8", a, b, o, None, None, None, None, 64)
    assert seen == [0]                                           # idle tuner: one launch, heuristic tile
    with autotune():
        lowp._launch("fp8", a, b, o, None, None, None, None, 64)
    assert seen[-1] == 128 and set(seen) == {0, 64, 128, 192, 256}
    n = len(seen)
    with autotune(tuning_buckets=(16, 256)):                     # scale tensors are tied to the live M: no synthetic buckets
        lowp._launch("fp8", a, b, o, None, None, None, None, 64)
    assert all(s_ in (0, 64, 128, 192, 256) for s_ in seen[n:]) and len(seen) - n <= 6
    seen.clear()
    a2, o2 = torch.zeros(1, 60, 64, dtype=torch.uint8), torch.zeros(1, 60, 512)
    lowp._launch("fp8", a2, b, o2, None, None, None, None, 64)   # same M bucket (64): cached choice, no profiling
    lowp._launch("fp8", a2, b, o2, None, None, None, None, 64, bn=256)   # an explicit tile always wins
    assert seen == [128, 256]


def test_shipped_and_env_configs_seed_the_singleton(tmp_path, monkeypatch):
    import flashinfer_b200.autotuner as at

    shipped, user = tmp_path / "shipped.json", tmp_path / "user.json"
    a = AutoTuner()
    a.profiling_cache[("op_s", "R", ((8,),), ())] = (0, 64, 0.1)
    a.save_configs(str(shipped))
    b = AutoTuner()
    b.profiling_cache[("op_u", "R", ((8,),), ())] = (0, 128, 0.1)
    b.save_configs(str(user))
    monkeypatch.setattr(at, "get_config_path", lambda is_module=False: str(shipped))
    monkeypatch.setenv("FLASHINFER_AUTOTUNER_CACHE", str(user))
    monkeypatch.setattr(AutoTuner, "_instance", None)
    t = AutoTuner.get()
    assert t.profiling_cache[("op_s", "R", ((8,),), ())][1] == 64 and t.profiling_cache[("op_u", "R", ((8,),), ())][1] == 128
    monkeypatch.setattr(AutoTuner, "_instance", None)          # the next get() builds a clean singleton for the other tests


def test_decode_linear_plan_is_a_tuner_client(tuner):
    """The flagship decode GEMM asks the tuner for its tile plan (BN, split-K cluster) only while tuning or with loaded configs;
    tactics are the admissible single-wave plans, the choice is cached per (M bucket, N, K, epilogue)."""
    from flashinfer_b200.gemm import decode_linear as dl

    plans = dl.admissible_plans(768, 4096)
    assert plans == [16 * 16 + 1, 32 * 16 + 2, 64 * 16 + 4, 128 * 16 + 8]          # narrow N: only deeper K splits reach more SMs
    assert all((p // 16) % (16 * (p % 16)) == 0 for p in dl.admissible_plans(6144, 4096))
    x, w = torch.zeros(48, 4096), torch.zeros(768, 4096)
    seen = []

    def launch(bn, s):
        seen.append((bn, s))
        time.sleep({(0, 0): 6e-3, (16, 1): 8e-3, (32, 2): 6e-3, (64, 4): 5e-3, (128, 8): 2e-4}[(bn, s)])

    assert dl._tuned_plan(x, w, 768, 4096, dl.EPI_PLAIN, launch) == (0, 0) and not seen     # idle tuner: planner default, no probing
    with autotune():
        assert dl._tuned_plan(x, w, 768, 4096, dl.EPI_PLAIN, launch) == (128, 8)
    assert set(seen) == {(0, 0), (16, 1), (32, 2), (64, 4), (128, 8)}
    seen.clear()
    assert dl._tuned_plan(torch.zeros(60, 4096), w, 768, 4096, dl.EPI_PLAIN, launch) == (128, 8) and not seen   # same M bucket: cached
    with autotune():
        dl._tuned_plan(x, w, 768, 4096, dl.EPI_GATED_SILU, launch)                         # another epilogue is another key
    assert seen


def test_example_tuned_config_feeds_decode_linear(tuner):
    """The shipped example (measured B200 sweep) loads and answers the keys decode_linear asks with."""
    import os
    import warnings

    import flashinfer_b200
    from flashinfer_b200.gemm import decode_linear as dl

    path = os.path.join(os.path.dirname(flashinfer_b200.__file__), "tuning_configs", "examples", "llama3_8b_decode_linear_NVIDIA_B200.json")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                     # measured on a B200, loaded on whatever runs the test
        assert tuner.load_configs(path) == 16
    calls = []
    launch = lambda bn, s: calls.append((bn, s))  # noqa: E731
    x = torch.zeros(64, 4096, dtype=torch.bfloat16)
    w_qkv_tp8 = torch.zeros(64, 768, 64, dtype=torch.bfloat16)       # BlockMajorK [K / 64, N, 64]
    assert dl._tuned_plan(x, w_qkv_tp8, 768, 4096, dl.EPI_ROPE_APPEND, launch) == (128, 8)
    w_gu = torch.zeros(64, 28672, 64, dtype=torch.bfloat16)
    assert dl._tuned_plan(x, w_gu, 28672, 4096, dl.EPI_GATED_SILU, launch) == (0, 0)      # planner default was the best plan
    assert not calls                                                                     # cache hits never launch anything
