mkdir -p gpurun_out
for k in coupling_fwd_tc_kernel coupling_bwd_p2_tc_kernel coupling_bwd_p1_tc_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 20 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 1 --warmup 1 --no-extras --precision bf16x3 > gpurun_out/ncu_$k.log 2>&1
  tail -2 gpurun_out/ncu_$k.log
done
ls -la gpurun_out/*.ncu-rep
