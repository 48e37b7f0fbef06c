"""Benchmark / test utilities.  Parity: reference flashinfer/testing/utils.py (bench_gpu_time :1546-1745,
FLOP / byte formulas :456-750, fp8 helpers :239-418)."""
from .utils import (  # noqa: F401
    attention_flops,
    attention_flops_with_actual_seq_lens,
    attention_tb_per_sec,
    attention_tb_per_sec_with_actual_seq_lens,
    attention_tflops_per_sec,
    attention_tflops_per_sec_with_actual_seq_lens,
    bench_gpu_time,
    bench_gpu_time_with_cuda_event,
    bench_gpu_time_with_cudagraph,
    dequantize_fp8,
    gemm_flops,
    get_l2_cache_size,
    measured_peaks,
    quantize_fp8,
    set_seed,
    sleep_after_kernel_run,
)

from .utils import bench_gpu_time_with_cupti  # noqa: F401,E402  (CUPTI activity records through torch.profiler)
from .utils import (  # noqa: F401,E402
    aggregate_gpu_time_across_ranks,
    bench_kineto,
    calculate_rotation_count,
    count_bytes,
    empty_suppress,
    per_block_cast_to_fp8,
    per_token_cast_to_fp8,
    suppress_stdout_stderr,
)
