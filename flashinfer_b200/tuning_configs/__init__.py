"""Shipped tuned-tactic files, one JSON per device name (``autotuner.get_config_path()``), in the format
``AutoTuner.save_configs`` writes.  Parity: reference flashinfer/tuning_configs/ (per-GPU python dict modules for the trtllm
fused-MoE tactics).  No file is shipped yet for B200: the launch heuristics in the native launchers are what the
published-shape numbers in ``profiles/`` were measured with; ``with autotune(cache=...)`` writes a file a deployment can drop
here: `AutoTuner.get()` loads it (then `$FLASHINFER_AUTOTUNER_CACHE`) when the singleton is first created."""
