mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 --kv-layout NHD > gpurun_out/r3_bench_nhd.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --kv-layout HND > gpurun_out/r3_bench_hnd.log 2>&1
timeout 300 env SHAPES=qkv VARIANTS="0:0:0:1,128:2:0:1,64:2:0:1,96:2:160:1" python tools/decode_linear_shapes.py > gpurun_out/r3_shapes_qkv.log 2>&1
timeout 300 env SHAPES=o,down VARIANTS="0:0:0:1,256:8:0:1,128:4:160:1,64:2:0:1" python tools/decode_linear_shapes.py > gpurun_out/r3_shapes_od.log 2>&1
timeout 300 env SHAPES=gate_up VARIANTS="0:0:0:1,256:1:0:1,224:1:0:1,192:1:0:1" python tools/decode_linear_shapes.py > gpurun_out/r3_shapes_gu.log 2>&1
timeout 300 env SHAPES=lm_head VARIANTS="256:1:0:1,256:1:104:1,240:1:0:1,256:1:104:0" python tools/decode_linear_shapes.py > gpurun_out/r3_shapes_lm.log 2>&1
timeout 300 python -m pytest tests/test_gpu_decode_linear.py -x -q 2>&1 | tail -3 > gpurun_out/r3_tests.log
for f in r3_bench_nhd r3_bench_hnd; do tail -1 gpurun_out/$f.log | cut -c1-200; done
for f in qkv od gu lm; do grep -v RESULT gpurun_out/r3_shapes_$f.log | tail -3 | cut -c1-400; done
cat gpurun_out/r3_tests.log
