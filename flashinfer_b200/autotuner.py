"""Tactic auto-tuner.  Parity: reference flashinfer/autotuner.py (TunableRunner / TuningConfig / AutoTuner.choose_one
:1045 / autotune() :465, JSON save/load of tuned configs).

B200-first differences: tactics are plain hashable values (e.g. the N-tile width of a tcgen05 GEMM), profiling uses
CUDA events with a cold L2 and optional CUDA-graph capture, and the cache key buckets dynamic dims to powers of two so a
serving engine tunes once per bucket.  Outside ``with autotune():`` the tuner only *looks up* cached choices (falls back
to the runner's heuristic tactic ``-1``), so the hot path never profiles.
"""
from __future__ import annotations

import contextlib
import json
import os
import threading
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch


def next_positive_power_of_2(x: int) -> int:
    return 1 if x < 1 else 1 << (int(x) - 1).bit_length()


def last_positive_power_of_2(x: int) -> int:
    n = next_positive_power_of_2(x)
    return n if n == x else n // 2


def get_power_of_2_num_tokens_buckets(max_num_tokens: int) -> Tuple[int, ...]:
    out, m = [], next_positive_power_of_2(max_num_tokens)
    while m >= 1:
        out.append(m)
        m //= 2
    return tuple(out)


@dataclass(frozen=True)
class DynamicTensorSpec:
    """``input_idx[i]`` / ``dim_idx[i]`` name tensor dims that vary at run time; ``gen_tuning_buckets`` lists the
    sizes to profile and ``map_to_tuning_buckets`` maps a live size to its bucket."""
    input_idx: Tuple[int, ...]
    dim_idx: Tuple[int, ...]
    gen_tuning_buckets: Tuple[int, ...] = ()
    map_to_tuning_buckets: Callable[[int], int] = next_positive_power_of_2
    tensor_initializers: Tuple[Callable, ...] = ()


@dataclass(frozen=True)
class ConstraintSpec:
    input_idx: int
    dim_idx: int
    infer_shape: Callable[[List[Tuple[int, ...]]], int]


@dataclass
class TuningConfig:
    dynamic_tensor_specs: Tuple[DynamicTensorSpec, ...] = ()
    constraint_specs: Tuple[ConstraintSpec, ...] = ()
    use_cuda_graph: bool = False
    use_cold_l2_cache: bool = True


class TunableRunner:
    """A kernel family with selectable tactics.  ``tactic == -1`` must always work (the built-in heuristic)."""

    def get_valid_tactics(self, inputs: List[torch.Tensor], profile: "OptimizationProfile") -> List[Any]:
        return [-1]

    def forward(self, inputs: List[torch.Tensor], tactic: Any = -1, do_preparation: bool = False, **kwargs):
        raise NotImplementedError

    def get_cache_key_extras(self, inputs: List[torch.Tensor]) -> Tuple:
        """Extra hashable state that distinguishes cached tactics beyond the input shapes (default: none)."""
        return ()

    def __call__(self, inputs, **kwargs):
        return self.forward(inputs, **kwargs)

    def __hash__(self):
        return hash(type(self).__name__)


@dataclass
class StaticDim:
    """A dimension with one value (reference autotuner.py:332)."""
    val: int

    def _opt(self) -> int:
        return self.val


@dataclass(unsafe_hash=True)
class DynamicDim:
    """Range of one dimension (reference autotuner.py:340)."""
    min: int
    opt: int
    max: int

    def _opt(self) -> int:
        return self.opt


@dataclass
class FakeTensor:
    """Shape-only stand-in used when profiles are enumerated without allocating (reference autotuner.py:375)."""
    dtype: torch.dtype
    device: torch.device
    shape: List[Any]


@dataclass
class AutoTunerStatistics:
    """Counters the tuner keeps (reference autotuner.py:697)."""
    cache_misses: int = 0
    cache_miss_config_collection: Dict[str, set] = field(default_factory=dict)
    failed_profiling_count: Dict[str, set] = field(default_factory=dict)
    tuned_op_total_configs: Dict[str, int] = field(default_factory=dict)
    tuned_op_successful_configs: Dict[str, int] = field(default_factory=dict)

    def __str__(self) -> str:
        return (f"Cache misses: {self.cache_misses}\nTuned ops: {dict(self.tuned_op_total_configs)}\n"
                f"Successful: {dict(self.tuned_op_successful_configs)}\n")


@dataclass
class OptimizationProfile:
    shapes: List[Tuple[int, ...]] = field(default_factory=list)
    tensor_initializers: List[Any] = field(default_factory=list)

    def key(self) -> Tuple:
        return tuple(self.shapes)

    def get_opt_shapes(self) -> Tuple:
        """Shapes with every Dim resolved to its tuning value (plain ints pass through)."""
        return tuple(tuple(d._opt() if hasattr(d, "_opt") else int(d) for d in shp) for shp in self.shapes)

    def get_hash_key(self) -> Tuple:
        return self.get_opt_shapes()


class AutoTuner:
    _instance: Optional["AutoTuner"] = None
    _lock = threading.Lock()

    def __init__(self, warmup: int = 3, repeat: int = 10) -> None:
        self.is_tuning_mode = False
        self.warmup, self.repeat = warmup, repeat
        self.profiling_cache: Dict[Tuple, Tuple[int, Any, float]] = {}
        self.stats = {"hits": 0, "misses": 0, "profiled": 0, "failed": 0}
        self._flush = None

    @classmethod
    def reset_statistics(self) -> None:
        self.stats = {"hits": 0, "misses": 0, "profiled": 0, "failed": 0}

    def get_effective_map_to_tuning_buckets(self, spec=None):
        """The bucket mapper of a dynamic-dimension spec, or the default next-power-of-two bucketing."""
        fn = getattr(spec, "map_to_tuning_buckets", None)
        return fn if fn is not None else (lambda x: 1 << max(0, int(x) - 1).bit_length())

    @classmethod
    def get(cls) -> "AutoTuner":
        with cls._lock:
            if cls._instance is None:
                cls._instance = AutoTuner()
            return cls._instance

    # ---- cache
    @staticmethod
    def _bucket_shapes(inputs: Sequence[Any], cfg: TuningConfig) -> Tuple:
        shapes = [tuple(t.shape) if isinstance(t, torch.Tensor) else () for t in inputs]
        shapes = [list(s) for s in shapes]
        for spec in cfg.dynamic_tensor_specs:
            for ii, dd in zip(spec.input_idx, spec.dim_idx):
                if ii < len(shapes) and dd < len(shapes[ii]):
                    shapes[ii][dd] = spec.map_to_tuning_buckets(shapes[ii][dd])
        for c in cfg.constraint_specs:
            if c.input_idx < len(shapes) and c.dim_idx < len(shapes[c.input_idx]):
                shapes[c.input_idx][c.dim_idx] = c.infer_shape([tuple(s) for s in shapes])
        return tuple(tuple(s) for s in shapes)

    def _key(self, op: str, runner: TunableRunner, shapes: Tuple, extras: Tuple = ()) -> Tuple:
        return (op, type(runner).__name__, shapes, extras)

    def search_cache(self, op: str, runners: Sequence[TunableRunner], shapes: Tuple, extras: Tuple = ()):
        for i, r in enumerate(runners):
            hit = self.profiling_cache.get(self._key(op, r, shapes, extras))
            if hit is not None:
                return True, hit[0], hit[1]
        return False, 0, -1

    def clear_cache(self) -> None:
        self.profiling_cache.clear()

    def save_configs(self, path: str) -> None:
        rows = [{"op": k[0], "runner": k[1], "shapes": [list(s) for s in k[2]], "extras": list(k[3]), "runner_id": v[0],
                 "tactic": v[1], "ms": v[2]} for k, v in self.profiling_cache.items()]
        with open(path, "w") as f:
            json.dump({"device": torch.cuda.get_device_name() if torch.cuda.is_available() else "cpu", "configs": rows}, f, indent=1)

    def load_configs(self, path: str) -> int:
        with open(path) as f:
            data = json.load(f)
        for r in data.get("configs", []):
            tactic = r["tactic"]
            if isinstance(tactic, list):
                tactic = tuple(tactic)
            key = (r["op"], r["runner"], tuple(tuple(s) for s in r["shapes"]), tuple(r.get("extras", [])))
            self.profiling_cache[key] = (r.get("runner_id", 0), tactic, r.get("ms", 0.0))
        return len(data.get("configs", []))

    # ---- profiling
    def _time(self, runner: TunableRunner, inputs, tactic, cfg: TuningConfig, **kwargs) -> float:
        if not torch.cuda.is_available():
            import time

            t0 = time.perf_counter()
            runner.forward(inputs, tactic=tactic, **kwargs)
            return (time.perf_counter() - t0) * 1e3
        for _ in range(self.warmup):
            runner.forward(inputs, tactic=tactic, **kwargs)
        torch.cuda.synchronize()
        if cfg.use_cold_l2_cache and self._flush is None:
            self._flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        graph = None
        if cfg.use_cuda_graph:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                runner.forward(inputs, tactic=tactic, **kwargs)
        ts = []
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(self.repeat):
            if cfg.use_cold_l2_cache:
                self._flush.zero_()
            s.record()
            if graph is not None:
                graph.replay()
            else:
                runner.forward(inputs, tactic=tactic, **kwargs)
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        ts.sort()
        return ts[len(ts) // 2]

    def choose_one(self, custom_op: str, runners: Sequence[TunableRunner], tuning_config: TuningConfig,
                   inputs: List[torch.Tensor], extras: Tuple = (), **kwargs) -> Tuple[TunableRunner, Any]:
        """Returns ``(runner, tactic)``: the cached best, or — inside ``autotune()`` — the freshly profiled best."""
        shapes = self._bucket_shapes(inputs, tuning_config)
        hit, rid, tactic = self.search_cache(custom_op, runners, shapes, extras)
        if hit:
            self.stats["hits"] += 1
            return runners[rid], tactic
        if not self.is_tuning_mode:
            self.stats["misses"] += 1
            return runners[0], -1
        best = (float("inf"), 0, -1)
        prof = OptimizationProfile([tuple(s) for s in shapes])
        for rid, r in enumerate(runners):
            for tac in r.get_valid_tactics(inputs, prof):
                try:
                    ms = self._time(r, inputs, tac, tuning_config, **kwargs)
                    self.stats["profiled"] += 1
                except Exception:  # noqa: BLE001 - a tactic that cannot run is simply not a candidate
                    self.stats["failed"] += 1
                    continue
                if ms < best[0]:
                    best = (ms, rid, tac)
        ms, rid, tac = best
        self.profiling_cache[self._key(custom_op, runners[rid], shapes, extras)] = (rid, tac, ms)
        return runners[rid], tac


@contextlib.contextmanager
def autotune(tune_mode: bool = True, cache_path: Optional[str] = None):
    """``with autotune():`` profiles every tunable op reached inside the block and caches the winners.
    ``cache_path`` (or ``$FLASHINFER_AUTOTUNER_CACHE``) loads existing choices first and saves on exit."""
    tuner = AutoTuner.get()
    cache_path = cache_path or os.environ.get("FLASHINFER_AUTOTUNER_CACHE")
    if cache_path and os.path.exists(cache_path):
        tuner.load_configs(cache_path)
    old = tuner.is_tuning_mode
    tuner.is_tuning_mode = tune_mode
    try:
        yield tuner
    finally:
        tuner.is_tuning_mode = old
        if cache_path and tune_mode:
            tuner.save_configs(cache_path)
