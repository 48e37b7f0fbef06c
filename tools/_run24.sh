mkdir -p gpurun_out
( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r24_bench8.log 2>&1
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 tools/tp_breakdown.py > gpurun_out/r24_tp8.log 2>&1
( timeout 200 python -m pytest tests/test_gpu_decode_linear.py tests/test_argmax_push.py -x -q -m gpu -k "8-two_shot or test_argmax_push and 8" 2>&1 | tail -4 ) > gpurun_out/r24_tests.log 2>&1
cat gpurun_out/r24_tests.log; grep -h '"metric"' gpurun_out/r24_bench8.log | cut -c1-6000; grep RESULT gpurun_out/r24_tp8.log
tail -2 gpurun_out/r24_bench8.log | cut -c1-300
