mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_allreduce_push.py -x -q -k "2" 2>&1 | tail -15 > gpurun_out/r11_push.log
timeout 600 python -m pytest tests/test_moe.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r11_moe.log
timeout 300 python tools/gemm_large_shapes.py > gpurun_out/r11_gemm.log 2>&1
timeout 600 python -m pytest tests/test_comm_multigpu.py tests/test_gemm_comm_fused.py -x -q -k "2" 2>&1 | tail -5 > gpurun_out/r11_comm.log
cat gpurun_out/r11_push.log; cat gpurun_out/r11_moe.log; grep RESULT gpurun_out/r11_gemm.log; cat gpurun_out/r11_comm.log
