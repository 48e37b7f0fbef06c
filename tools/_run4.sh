mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode_linear.py -x -q -k "tp_residual" 2>&1 | tail -5 > gpurun_out/r4_tests.log
P=29511
for mode in "" "--unfused"; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 20 --warmup 5 $mode > gpurun_out/r4_bench2$mode.log 2>&1
  P=$((P+1))
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r4_bench2_ref.log 2>&1
cat gpurun_out/r4_tests.log
tail -2 gpurun_out/r4_bench2.log | cut -c1-300; tail -2 gpurun_out/r4_bench2--unfused.log | cut -c1-300; tail -2 gpurun_out/r4_bench2_ref.log | cut -c1-1800
