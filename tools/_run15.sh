mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r15_tests.log 2>&1
TP=8 timeout 200 python tools/tp_breakdown.py > gpurun_out/r15_tp8_local.log 2>&1
TP=4 timeout 200 python tools/tp_breakdown.py > gpurun_out/r15_tp4_local.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r15_bench1.log 2>&1
FIB200_DL_SMEM_KB=100 timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r15_bench1_smem100.log 2>&1
FIB200_DL_SMEM_KB=110 FIB200_BENCH_KV_LAYOUT=HND timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r15_bench1_smem110_hnd.log 2>&1
cat gpurun_out/r15_tests.log; grep -h RESULT gpurun_out/r15_tp8_local.log gpurun_out/r15_tp4_local.log
for f in gpurun_out/r15_bench1*.log; do echo $f; grep -h '"metric"' $f | cut -c1-330; tail -2 $f | cut -c1-200; done
