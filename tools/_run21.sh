mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_prefill_variants.py tests/test_argmax_push.py tests/test_gpu_generic_attention.py tests/test_gpu_prefill_sampling.py tests/test_wrappers.py -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r21_tests.log 2>&1
cat gpurun_out/r21_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r21_bench1.log 2>&1
grep -h '"metric"' gpurun_out/r21_bench1.log | cut -c1-230; tail -2 gpurun_out/r21_bench1.log | cut -c1-200
