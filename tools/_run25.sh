mkdir -p gpurun_out
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/mc_probe.py > gpurun_out/r25_mc.log 2>&1
grep MCPROBE gpurun_out/r25_mc.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 benchmarks/extra_configs.py --config tp_gemm_rs > gpurun_out/r25_rs.log 2>&1
tail -3 gpurun_out/r25_rs.log | cut -c1-1500
