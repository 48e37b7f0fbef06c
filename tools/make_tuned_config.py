"""Turn a tools/dl_sweep.py log into an autotuner config file for the decode-linear tile plans.

    python tools/make_tuned_config.py gpurun_out/r14_sweep.log flashinfer_b200/tuning_configs/examples/llama3_8b_decode_linear_NVIDIA_B200.json

Every SWEEP line (one projection of a Llama-3-8B layer on a TP shard, measured on B200) becomes one entry keyed exactly like
`gemm.decode_linear._tuned_plan` keys its choices: op "decode_linear", runner "_PlanRunner", shapes of (x, BlockMajorK w) with the
token dimension bucketed, extras (epilogue, N, K, dtype, weight rank).  Load with FLASHINFER_AUTOTUNER_CACHE=<file> or
`with autotune(tune_mode=False, cache=<file>)`."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from flashinfer_b200.autotuner import AutoTuner, DynamicTensorSpec, TuningConfig
from flashinfer_b200.gemm import decode_linear as dl

EPI = {"qkv": dl.EPI_ROPE_APPEND, "gate_up": dl.EPI_GATED_SILU, "o": dl.EPI_RESIDUAL, "down": dl.EPI_RESIDUAL}


def main(log, out):
    tuner = AutoTuner()
    cfg = TuningConfig(dynamic_tensor_specs=(DynamicTensorSpec((0,), (0,)),), use_cold_l2_cache=True, synthesize_buckets=False)
    seen = set()
    for line in open(log):
        if not line.startswith("SWEEP "):
            continue
        r = json.loads(line[6:])
        key_id = (r["n"], r["k"], r["op"] if r["op"] in ("qkv", "gate_up") else "resid")
        if key_id in seen or not r["best"]:
            continue
        seen.add(key_id)
        t_best, bn, s = r["best"][0]
        tactic = bn * 16 + s if t_best < r["default_us"] * 0.98 else -1     # keep the built-in planner unless the sweep beat it by > 2 %
        n, k = r["n"], r["k"]
        shapes = tuner._bucket_shapes([torch.empty(64, k, device="meta"), torch.empty(k // 64, n, 64, device="meta")], cfg)
        extras = (int(EPI[r["op"]]), int(n), int(k), "torch.bfloat16", 3)
        tuner.profiling_cache[("decode_linear", "_PlanRunner", shapes, extras)] = (0, tactic, round(min(t_best, r["default_us"]) / 1e3, 6))
    if os.path.exists(out):
        os.remove(out)
    tuner.save_configs(out)
    data = json.load(open(out))  # the sweep ran on a B200 through gpurun; this script usually runs on the CPU-only build box
    data["metadata"].update({"device": "NVIDIA B200", "sm": "10.0", "source": os.path.basename(log)})
    json.dump(data, open(out, "w"), indent=1)
    print(f"{len(tuner.profiling_cache)} entries -> {out}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
