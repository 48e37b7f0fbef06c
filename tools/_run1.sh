mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode_linear.py -x -q 2>&1 | tail -15 > gpurun_out/r1_tests.log
timeout 400 env VARIANTS="0:0:104" python tools/decode_linear_shapes.py > gpurun_out/r1_shapes.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r1_bench_fused.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --unfused > gpurun_out/r1_bench_unfused.log 2>&1
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r1_bench_ref.log 2>&1
tail -5 gpurun_out/r1_tests.log; tail -8 gpurun_out/r1_shapes.log; tail -2 gpurun_out/r1_bench_fused.log; tail -2 gpurun_out/r1_bench_unfused.log; tail -3 gpurun_out/r1_bench_ref.log
