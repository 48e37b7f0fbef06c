mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_decode_linear.py tests/test_allreduce_push.py tests/test_gemm_comm_fused.py tests/test_moe_alltoall.py -x -q -k "8 and (tp_residual or push or gemm_comm or alltoall)" 2>&1 | tail -8 ) > gpurun_out/r12_tests.log 2>&1
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r12_bench8.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 tools/tp_breakdown.py > gpurun_out/r12_tp8.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 4 --steps 20 --warmup 5 --no-extras > gpurun_out/r12_bench4.log 2>&1
cat gpurun_out/r12_tests.log; grep -h '"metric"' gpurun_out/r12_bench8.log | cut -c1-5000; grep real gpurun_out/r12_bench8.log; grep RESULT gpurun_out/r12_tp8.log; grep -h '"metric"' gpurun_out/r12_bench4.log | cut -c1-400
tail -3 gpurun_out/r12_bench8.log | cut -c1-500
