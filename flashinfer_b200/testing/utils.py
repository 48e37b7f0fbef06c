"""GPU timing with cold-L2 discipline + roofline helpers.

Timing rules implemented here (the ones bench.py follows): warm-up first, CUDA events on the launching stream,
synchronise on both sides, the L2 is flushed between iterations (or inputs are rotated through buffers larger
than the L2), multi-rank numbers are the MAX over ranks.
"""
from __future__ import annotations

import contextlib
import json
import os
import time
from typing import Callable, List, Optional, Sequence, Tuple

import torch


def set_seed(seed: int = 0) -> None:
    import random

    import numpy as np

    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def get_l2_cache_size(device=None) -> int:
    if not torch.cuda.is_available():
        return 126 * 1024 * 1024
    return torch.cuda.get_device_properties(device or torch.cuda.current_device()).L2_cache_size


def measured_peaks() -> dict:
    """The driver-written roofline denominators (MEASURED_PEAKS.json) or the documented fall-backs."""
    here = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    path = os.path.join(here, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {"hbm_gbps": 6571.0, "bf16_tflops": 1640.0, "source": "fallback"}


def sleep_after_kernel_run(execution_time_ms: float) -> None:
    """Idle proportional to the kernel time so back-to-back measurements do not heat-throttle each other."""
    if execution_time_ms and execution_time_ms > 0:
        time.sleep(min(execution_time_ms / 200.0, 1.0))


class _L2Flusher:
    def __init__(self, device):
        self.buf = torch.empty(int(get_l2_cache_size(device) * 2), dtype=torch.uint8, device=device)

    def __call__(self):
        self.buf.zero_()


def bench_gpu_time_with_cuda_event(fn: Callable, dry_run_iters: Optional[int] = None, repeat_iters: Optional[int] = None,
                                   dry_run_time_ms: int = 25, repeat_time_ms: int = 100, l2_flush: bool = True,
                                   sleep_after_run: bool = False, input_args: Tuple = (), input_kwargs: Optional[dict] = None,
                                   cold_l2_cache: Optional[bool] = None, *, l2_flush_size_mb: Optional[int] = None,
                                   l2_flush_device: Optional[str] = None, aggregate_op=None) -> List[float]:
    """Per-iteration times (ms) measured with CUDA events; the L2 is flushed before every timed call.  ``l2_flush_size_mb`` /
    ``l2_flush_device`` (deprecated in the reference too) are accepted: the flush buffer is sized from the device's L2 here.
    ``aggregate_op`` is applied by :func:`bench_gpu_time` (cross-rank reduction of the samples)."""
    input_kwargs = input_kwargs or {}
    if cold_l2_cache is not None:
        l2_flush = cold_l2_cache
    dev = torch.cuda.current_device()
    flush = _L2Flusher(dev) if l2_flush else None
    call = lambda: fn(*input_args, **input_kwargs)  # noqa: E731
    # estimate
    call()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        call()
    e.record()
    torch.cuda.synchronize()
    est = max(s.elapsed_time(e) / 3, 1e-3)
    dry = dry_run_iters if dry_run_iters is not None else max(3, int(dry_run_time_ms / est))
    rep = repeat_iters if repeat_iters is not None else max(5, min(10000, int(repeat_time_ms / est)))
    for _ in range(max(dry, 3)):
        call()
    torch.cuda.synchronize()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(rep)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(rep)]
    for i in range(rep):
        if flush is not None:
            flush()
        starts[i].record()
        call()
        ends[i].record()
        if sleep_after_run:
            torch.cuda.synchronize()
            sleep_after_kernel_run(starts[i].elapsed_time(ends[i]))
    torch.cuda.synchronize()
    return [starts[i].elapsed_time(ends[i]) for i in range(rep)]


def bench_gpu_time_with_cudagraph(fn: Callable, dry_run_iters: Optional[int] = None, repeat_iters: Optional[int] = None,
                                  dry_run_time_ms: int = 25, repeat_time_ms: int = 100, num_iters_within_graph: int = 10,
                                  l2_flush: bool = True, sleep_after_run: bool = False, input_args: Tuple = (),
                                  input_kwargs: Optional[dict] = None, rotate_inputs: Optional[Sequence[Tuple]] = None,
                                  cold_l2_cache: Optional[bool] = None, *, l2_flush_size_mb: Optional[int] = None,
                                  l2_flush_device: Optional[str] = None, aggregate_op=None) -> List[float]:
    """Times a CUDA graph holding ``num_iters_within_graph`` calls (launch overhead amortised).  With
    ``rotate_inputs`` (a list of argument tuples whose total footprint exceeds the L2) every call inside the
    graph sees cold inputs; otherwise the L2 is flushed between graph replays only."""
    input_kwargs = input_kwargs or {}
    if cold_l2_cache is not None:
        l2_flush = cold_l2_cache
    dev = torch.cuda.current_device()
    flush = _L2Flusher(dev) if l2_flush else None
    argsets = list(rotate_inputs) if rotate_inputs else [tuple(input_args)]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for a in argsets[:2]:
            fn(*a, **input_kwargs)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(num_iters_within_graph):
            fn(*argsets[i % len(argsets)], **input_kwargs)
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    est = max(s.elapsed_time(e), 1e-3)
    dry = dry_run_iters if dry_run_iters is not None else max(3, int(dry_run_time_ms / est))
    rep = repeat_iters if repeat_iters is not None else max(5, min(2000, int(repeat_time_ms / est)))
    for _ in range(dry):
        g.replay()
    torch.cuda.synchronize()
    out = []
    for _ in range(rep):
        if flush is not None:
            flush()
        s.record()
        g.replay()
        e.record()
        torch.cuda.synchronize()
        t = s.elapsed_time(e) / num_iters_within_graph
        out.append(t)
        if sleep_after_run:
            sleep_after_kernel_run(t)
    return out


def bench_gpu_time_with_cupti(fn: Callable, dry_run_iters: Optional[int] = None, repeat_iters: Optional[int] = None,
                              dry_run_time_ms: int = 25, repeat_time_ms: int = 100, l2_flush: bool = True, sleep_after_run: bool = False,
                              input_args: Tuple = (), input_kwargs: Optional[dict] = None, cold_l2_cache: Optional[bool] = None,
                              use_cuda_graph: bool = False, *, l2_flush_size_mb: Optional[int] = None,
                              l2_flush_device: Optional[str] = None, aggregate_op=None) -> List[float]:
    """Per-iteration DEVICE time (ms) from CUPTI activity records: the sum of the durations of the kernels / copies an iteration
    launches, so launch gaps and host overhead are excluded (reference testing/utils.py:937, which reads the records through
    cupti-python; here through torch.profiler, whose Kineto backend records the same CUPTI activities).  Every iteration runs inside a
    ``record_function`` range; a device activity belongs to the iteration whose range contains its launch.  Falls back to
    :func:`bench_gpu_time_with_cuda_event` when no device activity can be attributed (profiler unavailable, graph replays)."""
    input_kwargs = input_kwargs or {}
    if cold_l2_cache is not None:
        l2_flush = cold_l2_cache
    fallback = lambda: bench_gpu_time_with_cuda_event(fn, dry_run_iters, repeat_iters, dry_run_time_ms, repeat_time_ms, l2_flush,  # noqa: E731
                                                      sleep_after_run, input_args, input_kwargs)
    if use_cuda_graph:
        return fallback()
    try:
        from torch.profiler import ProfilerActivity, profile, record_function
    except ImportError:
        return fallback()
    call = lambda: fn(*input_args, **input_kwargs)  # noqa: E731
    flush = _L2Flusher(torch.cuda.current_device()) if l2_flush else None
    call()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        call()
    e.record()
    torch.cuda.synchronize()
    est = max(s.elapsed_time(e) / 3, 1e-3)
    dry = dry_run_iters if dry_run_iters is not None else max(3, int(dry_run_time_ms / est))
    rep = repeat_iters if repeat_iters is not None else max(5, min(2000, int(repeat_time_ms / est)))
    for _ in range(max(dry, 3)):
        call()
    torch.cuda.synchronize()
    tag = "fib200_bench_iteration"
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(rep):
            if flush is not None:
                flush()
            with record_function(tag):
                call()
            if sleep_after_run:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
    events = prof.events()
    ranges = sorted((ev.time_range.start, ev.time_range.end) for ev in events if ev.name == tag)
    if len(ranges) != rep:
        return fallback()
    totals = [0.0] * rep
    starts = [r[0] for r in ranges]
    import bisect

    attributed = 0
    for ev in events:
        kernels = getattr(ev, "kernels", None)
        if not kernels or ev.name == tag:
            continue
        i = bisect.bisect_right(starts, ev.time_range.start) - 1
        if i < 0 or ev.time_range.start > ranges[i][1]:
            continue                                          # launched outside the timed ranges (the L2 flush)
        totals[i] += sum(k.duration for k in kernels)         # microseconds
        attributed += len(kernels)
    if attributed == 0:
        return fallback()
    return [t / 1e3 for t in totals]


def bench_gpu_time(fn: Callable, dry_run_iters: Optional[int] = None, repeat_iters: Optional[int] = None,
                   dry_run_time_ms: int = 25, repeat_time_ms: int = 100, l2_flush: bool = True, use_cuda_graph: bool = False,
                   num_iters_within_graph: int = 10, sleep_after_run: bool = False, enable_cupti: bool = False,
                   input_args: Tuple = (), input_kwargs: Optional[dict] = None, cold_l2_cache: Optional[bool] = None,
                   aggregate_op: Optional[str] = None, group=None, *, l2_flush_size_mb: Optional[int] = None,
                   l2_flush_device: Optional[str] = None) -> List[float]:
    """Unified entry (reference testing/utils.py:1546).  ``enable_cupti`` measures device time from CUPTI activity records
    (:func:`bench_gpu_time_with_cupti`).  ``aggregate_op='max'`` reduces every sample over the ranks of ``group``."""
    if enable_cupti and not use_cuda_graph:
        times = bench_gpu_time_with_cupti(fn, dry_run_iters, repeat_iters, dry_run_time_ms, repeat_time_ms, l2_flush, sleep_after_run,
                                          input_args, input_kwargs, cold_l2_cache)
    elif use_cuda_graph:
        times = bench_gpu_time_with_cudagraph(fn, dry_run_iters, repeat_iters, dry_run_time_ms, repeat_time_ms,
                                              num_iters_within_graph, l2_flush, sleep_after_run, input_args, input_kwargs,
                                              None, cold_l2_cache)
    else:
        times = bench_gpu_time_with_cuda_event(fn, dry_run_iters, repeat_iters, dry_run_time_ms, repeat_time_ms, l2_flush,
                                               sleep_after_run, input_args, input_kwargs, cold_l2_cache)
    if aggregate_op and torch.distributed.is_available() and torch.distributed.is_initialized():
        t = torch.tensor(times, dtype=torch.float64, device="cuda")
        n = torch.tensor([t.numel()], device="cuda")
        torch.distributed.all_reduce(n, op=torch.distributed.ReduceOp.MIN, group=group)
        t = t[: int(n)]
        op = {"max": torch.distributed.ReduceOp.MAX, "min": torch.distributed.ReduceOp.MIN,
              "sum": torch.distributed.ReduceOp.SUM}[getattr(aggregate_op, "__name__", aggregate_op)]      # "max" or the builtin max (reference)
        torch.distributed.all_reduce(t, op=op, group=group)
        times = t.tolist()
    return times


# ------------------------------------------------------------------ roofline formulas
def attention_flops(batch_size: int, qo_seqlen: int, kv_seqlen: int, head_dim_qk: int, head_dim_vo: int, num_qo_heads: int,
                    causal: bool) -> float:
    if causal:
        # rows see kv_len - qo_len + i + 1 keys
        pairs = qo_seqlen * kv_seqlen - qo_seqlen * (qo_seqlen - 1) / 2 if kv_seqlen >= qo_seqlen else kv_seqlen * (kv_seqlen + 1) / 2
    else:
        pairs = qo_seqlen * kv_seqlen
    return 2.0 * batch_size * num_qo_heads * pairs * (head_dim_qk + head_dim_vo)


def attention_flops_with_actual_seq_lens(actual_seq_lens_q, actual_seq_lens_kv, head_dim_qk: int, head_dim_vo: int,
                                         num_qo_heads: int, causal: bool) -> float:
    q = torch.as_tensor(actual_seq_lens_q).double().flatten()
    k = torch.as_tensor(actual_seq_lens_kv).double().flatten()
    if causal:
        pairs = torch.where(k >= q, q * k - q * (q - 1) / 2, k * (k + 1) / 2)
    else:
        pairs = q * k
    return float(2.0 * num_qo_heads * pairs.sum() * (head_dim_qk + head_dim_vo))


def attention_tflops_per_sec(batch_size, qo_seqlen, kv_seqlen, head_dim_qk, head_dim_vo, num_qo_heads, causal, time_ms) -> float:
    return attention_flops(batch_size, qo_seqlen, kv_seqlen, head_dim_qk, head_dim_vo, num_qo_heads, causal) / time_ms / 1e9


def attention_tflops_per_sec_with_actual_seq_lens(actual_seq_lens_q, actual_seq_lens_kv, head_dim_qk, head_dim_vo,
                                                  num_qo_heads, causal, time_ms) -> float:
    return attention_flops_with_actual_seq_lens(actual_seq_lens_q, actual_seq_lens_kv, head_dim_qk, head_dim_vo,
                                                num_qo_heads, causal) / time_ms / 1e9


def _bytes_of(dtype: torch.dtype) -> float:
    return torch.empty((), dtype=dtype).element_size()


def attention_tb_per_sec(batch_size, qo_seqlen, kv_seqlen, head_dim_qk, head_dim_vo, num_qo_heads, num_kv_heads, time_ms,
                         q_dtype=torch.bfloat16, kv_dtype=torch.bfloat16, o_dtype=torch.bfloat16) -> float:
    q = batch_size * qo_seqlen * num_qo_heads * head_dim_qk * _bytes_of(q_dtype)
    k = batch_size * kv_seqlen * num_kv_heads * head_dim_qk * _bytes_of(kv_dtype)
    v = batch_size * kv_seqlen * num_kv_heads * head_dim_vo * _bytes_of(kv_dtype)
    o = batch_size * qo_seqlen * num_qo_heads * head_dim_vo * _bytes_of(o_dtype)
    return (q + k + v + o) / time_ms / 1e9


def attention_tb_per_sec_with_actual_seq_lens(actual_seq_lens_q, actual_seq_lens_kv, head_dim_qk, head_dim_vo, num_qo_heads,
                                              num_kv_heads, time_ms, q_dtype=torch.bfloat16, kv_dtype=torch.bfloat16,
                                              o_dtype=torch.bfloat16) -> float:
    sq = float(torch.as_tensor(actual_seq_lens_q).double().sum())
    sk = float(torch.as_tensor(actual_seq_lens_kv).double().sum())
    b = sq * num_qo_heads * (head_dim_qk * _bytes_of(q_dtype) + head_dim_vo * _bytes_of(o_dtype))
    b += sk * num_kv_heads * (head_dim_qk + head_dim_vo) * _bytes_of(kv_dtype)
    return b / time_ms / 1e9


def gemm_flops(m: int, n: int, k: int, batch: int = 1) -> float:
    return 2.0 * batch * m * n * k


# ------------------------------------------------------------------ fp8 helpers
def quantize_fp8(x: torch.Tensor, scale_shape: Tuple[int, ...], tile_shape: Tuple[int, ...], scale_major_mode: str = "K"):
    """Tile-wise fp8 (e4m3) quantisation: returns (x_fp8, fp32 scales of ``scale_shape``); ``tile_shape`` is the block of
    elements sharing one scale (e.g. (1, 128) for activations, (128, 128) for weights)."""
    assert x.dim() == len(tile_shape)
    fmax = torch.finfo(torch.float8_e4m3fn).max
    pads = []
    shape = []
    for d, t in zip(x.shape, tile_shape):
        pads.append((d + t - 1) // t * t)
        shape += [pads[-1] // t, t]
    xp = torch.zeros(pads, dtype=torch.float32, device=x.device)
    xp[tuple(slice(0, d) for d in x.shape)] = x.float()
    blk = xp.view(shape)
    red = tuple(range(1, 2 * x.dim(), 2))
    amax = blk.abs().amax(red).clamp(min=1e-8)
    scale = amax / fmax
    idx = []
    for i in range(x.dim()):
        idx += [slice(None), None]
    q = (blk / scale[tuple(idx)]).view(pads)[tuple(slice(0, d) for d in x.shape)].clamp(-fmax, fmax).to(torch.float8_e4m3fn)
    if scale_major_mode == "MN" and scale.dim() == 2:
        scale = scale.t().contiguous().t()
    return q, scale


def dequantize_fp8(x_fp8: torch.Tensor, scale: torch.Tensor, scale_major_mode: str = "K",
                   tile_shape: Optional[Tuple[int, ...]] = None) -> torch.Tensor:
    s = scale.float()
    for d in range(x_fp8.dim()):
        if tile_shape is not None:
            rep = tile_shape[d]
        else:
            rep = (x_fp8.shape[d] + s.shape[d] - 1) // s.shape[d]
            if s.shape[d] > 1:  # tiles are powers of two in practice; the last one may be ragged
                rep = 1 << (rep - 1).bit_length()
        s = s.repeat_interleave(rep, d)
    s = s[tuple(slice(0, d) for d in x_fp8.shape)]
    return x_fp8.float() * s


# ------------------------------------------------------------------ DeepGEMM-style casts, byte counting, misc (reference testing/utils.py)
def _ceil_to_ue8m0(x: torch.Tensor) -> torch.Tensor:
    return torch.pow(2.0, torch.ceil(torch.log2(x.abs())))


def per_token_cast_to_fp8(x: torch.Tensor):
    """``[m, n]`` -> (e4m3 ``[m, n]``, power-of-two scales ``[m, n / 128]``), one scale per 1 x 128 group."""
    assert x.dim() == 2 and x.size(1) % 128 == 0
    m, n = x.shape
    xv = x.view(m, -1, 128)
    sf = _ceil_to_ue8m0(xv.abs().float().amax(dim=2).view(m, -1).clamp(1e-4) / 448.0)
    return (xv * (1.0 / sf.unsqueeze(2))).to(torch.float8_e4m3fn).view(m, n), sf


def per_block_cast_to_fp8(x: torch.Tensor):
    """``[m, n]`` -> (e4m3 ``[m, n]``, power-of-two scales ``[ceil(m/128), ceil(n/128)]``), one scale per 128 x 128 block."""
    assert x.dim() == 2
    m, n = x.shape
    mp, np_ = (m + 127) // 128 * 128, (n + 127) // 128 * 128
    xp = torch.zeros(mp, np_, dtype=x.dtype, device=x.device)
    xp[:m, :n] = x
    xv = xp.view(-1, 128, np_ // 128, 128)
    sf = _ceil_to_ue8m0(xv.abs().float().amax(dim=(1, 3), keepdim=True).clamp(1e-4) / 448.0)
    return (xv * (1.0 / sf)).to(torch.float8_e4m3fn).view_as(xp)[:m, :n].contiguous(), sf.view(xv.size(0), xv.size(2))


def count_bytes(*tensors) -> int:
    total = 0
    for t in tensors:
        if isinstance(t, (tuple, list)):
            total += count_bytes(*t)
        elif t is not None:
            total += t.numel() * t.element_size()
    return total


def calculate_rotation_count(tensors, device=None, min_rotations: int = 2) -> int:
    """How many rotating copies of ``tensors`` keep the L2 cold between benchmark iterations (1 if they dwarf the L2)."""
    nbytes = count_bytes(*tensors)
    l2 = 126 << 20
    if torch.cuda.is_available():
        l2 = torch.cuda.get_device_properties(device if device is not None else torch.cuda.current_device()).L2_cache_size or l2
    if nbytes >= 5 * l2:
        return 1
    return min(128, max(min_rotations, int(-(-2 * l2 // max(nbytes, 1))) + 1))


def aggregate_gpu_time_across_ranks(x, op):
    """Combine a per-rank time (or list of times) over the process group with ``op`` (e.g. ``max``)."""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size() > 1:
        allx = [None] * dist.get_world_size()
        dist.all_gather_object(allx, x)
        if isinstance(x, list):
            return [op(v) for v in zip(*allx)]
        return op(allx)
    return x


class empty_suppress:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class suppress_stdout_stderr:
    """Silence fd 1 / 2 inside the block (native libraries included)."""

    def __enter__(self):
        import os
        import sys

        sys.stdout.flush()
        sys.stderr.flush()
        self._null = os.open(os.devnull, os.O_WRONLY)
        self._saved = (os.dup(1), os.dup(2))
        os.dup2(self._null, 1)
        os.dup2(self._null, 2)
        return self

    def __exit__(self, *exc):
        import os

        os.dup2(self._saved[0], 1)
        os.dup2(self._saved[1], 2)
        for fd in (self._null, *self._saved):
            os.close(fd)
        return False


def bench_kineto(fn, kernel_names, num_tests: int = 30, suppress_kineto_output: bool = False, trace_path=None, barrier_comm_profiling=False,
                 flush_l2: bool = True):
    """Average device time (seconds) of the kernels whose name contains ``kernel_names`` (str or tuple), via torch.profiler."""
    from torch.profiler import ProfilerActivity, profile

    names = (kernel_names,) if isinstance(kernel_names, str) else tuple(kernel_names)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if flush_l2 else None
    fn()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(num_tests):
            if flush is not None:
                flush.zero_()
            fn()
        torch.cuda.synchronize()
    out = []
    for name in names:
        evs = [e for e in prof.key_averages() if name in e.key]
        tot = sum(getattr(e, "device_time_total", getattr(e, "cuda_time_total", 0.0)) for e in evs)
        out.append(tot / num_tests / 1e6)
    return out[0] if isinstance(kernel_names, str) else tuple(out)


class empty_suppress:
    """No-op stand-in for :class:`suppress_stdout_stderr` (reference testing/utils.py:1701)."""

    def __enter__(self):
        return self

    def __exit__(self, *_):
        return False


class suppress_stdout_stderr(contextlib.ExitStack):
    """Silence both the Python-level streams and file descriptors 1 / 2 (what native libraries print to) inside the block
    (reference testing/utils.py:1709)."""

    def __enter__(self):
        super().__enter__()
        import sys

        sink = self.enter_context(open(os.devnull, "w"))
        for stream in (sys.stdout, sys.stderr):
            try:
                fd = stream.fileno()
            except (AttributeError, OSError, ValueError):
                continue                                      # a captured / replaced stream without a descriptor
            stream.flush()
            saved = os.dup(fd)
            self.callback(os.close, saved)
            self.callback(os.dup2, saved, fd)
            os.dup2(sink.fileno(), fd)
        self.enter_context(contextlib.redirect_stdout(sink))
        self.enter_context(contextlib.redirect_stderr(sink))
        return self
