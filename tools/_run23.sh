mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_decode_linear.py tests/test_argmax_push.py tests/test_legacy_alltoall.py tests/test_gemm_comm_fused.py tests/test_allreduce_push.py tests/test_moe_alltoall.py -x -q -m gpu -k "4" 2>&1 | tail -6 ) > gpurun_out/r23_tests.log 2>&1
cat gpurun_out/r23_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29563 tools/tp_breakdown.py > gpurun_out/r23_tp4.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29564 bench.py --gpus 4 --steps 20 --warmup 5 --no-extras > gpurun_out/r23_bench4.log 2>&1
FIB200_DL_AR_ALGO=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29565 bench.py --gpus 4 --steps 20 --warmup 5 --no-extras > gpurun_out/r23_bench4_oneshot.log 2>&1
grep -h '"metric"' gpurun_out/r23_bench4.log gpurun_out/r23_bench4_oneshot.log | cut -c1-230
grep -h RESULT gpurun_out/r23_tp4.log
