mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_quant_topk.py -x -q -m gpu 2>&1 | tail -6 ) > gpurun_out/r20_tests.log 2>&1
cat gpurun_out/r20_tests.log
for t in dlinear_gate_up dlinear_down; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:dlinear -s 3 -c 1 -f -o gpurun_out/$t python tools/prof_targets.py $t > gpurun_out/r20_ncu_$t.log 2>&1
  tail -2 gpurun_out/r20_ncu_$t.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dlinear|decode_paged|decode_prep|argmax|reduce" -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --eager-steps 2 --no-extras > gpurun_out/r20_launches.log 2>&1
tail -3 gpurun_out/r20_launches.log; wc -l gpurun_out/launches_r2.csv
