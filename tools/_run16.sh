mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gemm_comm_fused.py tests/test_argmax_push.py tests/test_gpu_decode_linear.py -x -q -k "2" 2>&1 | tail -6 ) > gpurun_out/r16_tests.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29564 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r16_bench2.log 2>&1
cat gpurun_out/r16_tests.log; grep -h '"metric"' gpurun_out/r16_bench2.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['n_gpus'], d['ms_per_step'], d['e2e'], d['gpu_launches_native_per_step'])
    print(json.dumps(d['extra'].get('tp_gemm_rs')))
"
tail -3 gpurun_out/r16_bench2.log | cut -c1-300
