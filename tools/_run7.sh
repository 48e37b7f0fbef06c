mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r7_tests.log 2>&1
timeout 300 python tools/profile_decode_linear.py gpurun_out > gpurun_out/r7_profile.log 2>&1
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r7_bench.log 2>&1
( time timeout 1500 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r7_bench_ref.log 2>&1
tail -8 gpurun_out/r7_tests.log; grep RESULT gpurun_out/r7_profile.log | cut -c1-1500; tail -5 gpurun_out/r7_bench.log | cut -c1-3000; tail -5 gpurun_out/r7_bench_ref.log | cut -c1-3500
