#!/bin/bash
# tests for multi-tensor Adam + ncu: launch list of one bench step and full captures of the 4 main kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --precision bf16x3 > gpurun_out/ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/r01_launches.csv | head -30
for k in coupling_fwd_train_tc2_kernel coupling_bwd_p1_tc2_kernel coupling_bwd_p2_tc2_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 70 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 1 --warmup 3 --no-extras --precision bf16x3 > gpurun_out/ncu_$k.log 2>&1
  tail -2 gpurun_out/ncu_$k.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decoder_eval_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_decoder_eval_tc_kernel python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_eval.log 2>&1
tail -2 gpurun_out/ncu_eval.log
ls -la gpurun_out/*.ncu-rep
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01_full.json 2> gpurun_out/bench_r01_full.log
cat gpurun_out/bench_r01_full.json
