mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode_linear.py -x -q 2>&1 | tail -15 > gpurun_out/r2_tests.log
timeout 500 env VARIANTS="0:0:0:1,128:4:0:0,128:4:0:1,192:4:0:0,192:4:0:1" python tools/decode_linear_shapes.py > gpurun_out/r2_shapes.log 2>&1
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2_bench_ref.log 2>&1
( time timeout 900 python bench.py --impl reference --ref-src wheel --steps 20 --warmup 5 ) > gpurun_out/r2_bench_ref_wheel.log 2>&1
tail -5 gpurun_out/r2_tests.log; tail -6 gpurun_out/r2_shapes.log | cut -c1-400; tail -5 gpurun_out/r2_bench_ref.log | cut -c1-1500; tail -5 gpurun_out/r2_bench_ref_wheel.log | cut -c1-1500
