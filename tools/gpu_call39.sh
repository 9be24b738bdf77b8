#!/bin/bash
# round-1 final evidence: smoke, full GPU test suite, bench (both arms), launch list, full ncu captures
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.log; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r01_launches_final.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/r01_launches_final.csv | head -14
for k in coupling_fwd_train_tc2_kernel coupling_bwd_p1_tc2_kernel coupling_bwd_p2_tc4_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 70 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pointnet_eval_kernel|pairwise_cd_kernel" -c 2 -f -o gpurun_out/prof_pointnet_chamfer python bench.py --steps 1 --warmup 3 --sweep-clouds 0 > gpurun_out/ncu_extras.log 2>&1
tail -1 gpurun_out/ncu_extras.log
