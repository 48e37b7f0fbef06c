mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_legacy_alltoall.py tests/test_sparse_mla.py tests/test_gpu_attention_gemm.py tests/test_gpu_mla.py -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r19_tests.log 2>&1
cat gpurun_out/r19_tests.log
