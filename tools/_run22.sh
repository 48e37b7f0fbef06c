mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_decode_linear.py -x -q -m gpu 2>&1 | tail -5 ) > gpurun_out/r22_tests.log 2>&1
cat gpurun_out/r22_tests.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r22_bench1.log 2>&1 &
P1=$!
wait $P1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 tools/tp_breakdown.py > gpurun_out/r22_tp2.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29564 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/r22_bench2.log 2>&1
TP=1 timeout 200 python tools/tp_breakdown.py > gpurun_out/r22_tp1_local.log 2>&1
grep -h '"metric"' gpurun_out/r22_bench1.log gpurun_out/r22_bench2.log | cut -c1-230
grep -h RESULT gpurun_out/r22_tp2.log gpurun_out/r22_tp1_local.log
