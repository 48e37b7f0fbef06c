mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_comm_fused.py -x -q -k "2" 2>&1 | tail -5 > gpurun_out/r9_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 benchmarks/extra_configs.py --impl ours --config tp_gemm_rs > gpurun_out/r9_rs.log 2>&1
cat gpurun_out/r9_tests.log; grep extra gpurun_out/r9_rs.log | cut -c1-1500 || tail -20 gpurun_out/r9_rs.log
