mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_attention_gemm.py tests/test_wrappers.py tests/test_gpu_decode_linear.py tests/test_gpu_generic_attention.py -x -q -m gpu 2>&1 | tail -5 ) > gpurun_out/r18_tests.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r18_bench1.log 2>&1
FIB200_KV_PREFETCH=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r18_bench1_nopf.log 2>&1
TP=8 timeout 200 python tools/tp_breakdown.py > gpurun_out/r18_tp8_local.log 2>&1
cat gpurun_out/r18_tests.log
for f in gpurun_out/r18_bench1*.log; do echo $f; grep -h '"metric"' $f | cut -c1-230; tail -2 $f | cut -c1-200; done
grep RESULT gpurun_out/r18_tp8_local.log
