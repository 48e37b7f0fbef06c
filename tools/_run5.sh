mkdir -p gpurun_out
timeout 300 env TP=8 python tools/tp_breakdown.py > gpurun_out/r5_tp8_local.log 2>&1
timeout 300 env TP=2 python tools/tp_breakdown.py > gpurun_out/r5_tp2_local.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/tp_breakdown.py > gpurun_out/r5_tp2.log 2>&1
grep RESULT gpurun_out/r5_tp8_local.log; grep RESULT gpurun_out/r5_tp2_local.log; grep RESULT gpurun_out/r5_tp2.log || tail -20 gpurun_out/r5_tp2.log
