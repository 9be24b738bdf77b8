#!/bin/bash
# re-entry baseline: full GPU test suite, smoke, both bench arms, launch list, full ncu captures of the main kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_full.json 2> gpurun_out/bench_r01_full.log; echo "bench rc=$?"
cat gpurun_out/bench_r01_full.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_reference.json 2> gpurun_out/bench_r01_reference.log; echo "ref rc=$?"
cat gpurun_out/bench_r01_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/r01_launches.csv | head -30
for k in coupling_fwd_train_tc2_kernel coupling_bwd_p1_tc2_kernel coupling_bwd_p2_tc2_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 70 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"decoder_eval_tc_kernel|pointnet_eval_kernel|pairwise_cd_kernel" -c 6 -f -o gpurun_out/prof_extras python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_extras.log 2>&1
tail -1 gpurun_out/ncu_extras.log
ls -la gpurun_out/*.ncu-rep
