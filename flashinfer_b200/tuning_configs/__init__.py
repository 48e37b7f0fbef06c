"""Shipped tuned-tactic files, one JSON per device name (``autotuner.get_config_path()``), in the format
``AutoTuner.save_configs`` writes.  Parity: reference flashinfer/tuning_configs/ (per-GPU python dict modules for the trtllm
fused-MoE tactics).  No file is shipped yet for B200: the launch heuristics in the native launchers are what the
published-shape numbers in ``profiles/`` were measured with; ``with autotune(cache=...)`` writes a file a deployment can drop
here: `AutoTuner.get()` loads it (then `$FLASHINFER_AUTOTUNER_CACHE`) when the singleton is first created.

``examples/llama3_8b_decode_linear_NVIDIA_B200.json``: the tile plans (BN, split-K cluster) of the decode linear for the Llama-3-8B
projections at TP 1 / 2 / 4 / 8, generated from the measured sweep (``tools/dl_sweep.py`` -> ``tools/make_tuned_config.py``;
``-1`` = the built-in planner was within 2 % of the best plan).  It is an example of the format and is NOT loaded automatically:
``FLASHINFER_AUTOTUNER_CACHE=<path>`` or ``with autotune(tune_mode=False, cache=<path>)``."""
