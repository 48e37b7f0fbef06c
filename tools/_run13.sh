mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_decode_linear.py -x -q -k "tp_residual and 2-" 2>&1 | tail -5 ) > gpurun_out/r13_tests.log 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 tools/tp_breakdown.py > gpurun_out/r13_tp2.log 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29564 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/r13_bench2.log 2>&1
FIB200_DL_AR_ALGO=2 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29565 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/r13_bench2_twoshot.log 2>&1
cat gpurun_out/r13_tests.log; grep RESULT gpurun_out/r13_tp2.log; grep -h '"metric"' gpurun_out/r13_bench2.log gpurun_out/r13_bench2_twoshot.log | cut -c1-300
tail -3 gpurun_out/r13_tp2.log | cut -c1-300
