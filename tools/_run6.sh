mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode_linear.py -x -q -k "tp_residual" 2>&1 | tail -5 > gpurun_out/r6_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/tp_breakdown.py > gpurun_out/r6_tp2.log 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r6_bench2.log 2>&1
cat gpurun_out/r6_tests.log; grep RESULT gpurun_out/r6_tp2.log || tail -20 gpurun_out/r6_tp2.log; tail -1 gpurun_out/r6_bench2.log | cut -c1-300
