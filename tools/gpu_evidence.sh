#!/bin/bash
# One-call evidence pass on a B200 box (run through gpurun from the dev container):
#   gpurun --timeout 2400 -- 'bash tools/gpu_evidence.sh'
# smoke, full GPU test suite, both bench arms, ncu launch list and full captures of the per-layer kernels.
# Outputs land in gpurun_out/ (scratch); copy what should be kept into profiles/ (tools/ncu_summary.py,
# tools/summarize_launches.py reduce the captures).
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.log; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; echo "reference arm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt; head -12 gpurun_out/launches_summary.txt
for k in coupling_fwd_train_tc2_kernel coupling_bwd_p1_tc2_kernel coupling_bwd_p2_tc4_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 70 -c 1 -f -o gpurun_out/prof_$k \
    python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
