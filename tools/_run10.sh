mkdir -p gpurun_out
timeout 300 python tools/gemm_large_shapes.py > gpurun_out/r10_gemm.log 2>&1
timeout 600 python -m pytest tests/test_gpu_attention_gemm.py tests/test_gpu_lowp_gemm.py -x -q 2>&1 | tail -5 > gpurun_out/r10_tests.log
grep RESULT gpurun_out/r10_gemm.log || tail -20 gpurun_out/r10_gemm.log; cat gpurun_out/r10_tests.log
