mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r8_bench2.log 2>&1
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r8_bench2_ref.log 2>&1
timeout 300 python tools/profile_decode_linear.py gpurun_out > gpurun_out/r8_profile.log 2>&1
grep -h '"metric"' gpurun_out/r8_bench2.log | cut -c1-4000; grep -h '"metric"' gpurun_out/r8_bench2_ref.log | cut -c1-4000; grep real gpurun_out/r8_bench2.log gpurun_out/r8_bench2_ref.log; grep RESULT gpurun_out/r8_profile.log | cut -c1-2000
tail -5 gpurun_out/r8_bench2.log | cut -c1-600
