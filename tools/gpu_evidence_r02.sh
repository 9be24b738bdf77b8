#!/bin/bash
# Round-2 evidence pass on one B200 (run through gpurun from the dev container):
#   gpurun --timeout 2400 -- 'bash tools/gpu_evidence_r02.sh'
# ncu launch list of the bench command + full captures of the hot kernels; outputs in gpurun_out/ (scratch) - reduce
# them with tools/ncu_summary.py / tools/summarize_launches.py and commit the summaries under profiles/r02_*.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r02_ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.txt; head -16 gpurun_out/r02_launches_summary.txt
for k in coupling_fwd_train_tc2_kernel coupling_bwd_p1_tc2_kernel coupling_bwd_p2_tc4_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 70 -c 1 -f -o gpurun_out/r02_prof_$k \
    python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/r02_ncu_$k.log 2>&1
  tail -1 gpurun_out/r02_ncu_$k.log
  python tools/ncu_summary.py gpurun_out/r02_prof_$k.ncu-rep > gpurun_out/r02_ncu_${k}_summary.txt; rm -f gpurun_out/r02_prof_$k.ncu-rep
done
# merged backward (opt-in form), sampling kernel, Chamfer all-pairs (one-evaluation kernel), fused EMD, score reduction
timeout 600 ncu --set full --clock-control none --import-source on -k regex:coupling_bwd_merged_kernel -s 70 -c 1 -f -o gpurun_out/r02_prof_coupling_bwd_merged_kernel \
  python bench.py --steps 1 --warmup 3 --no-extras --lib-option 5=1 > gpurun_out/r02_ncu_merged.log 2>&1; tail -1 gpurun_out/r02_ncu_merged.log
python tools/ncu_summary.py gpurun_out/r02_prof_coupling_bwd_merged_kernel.ncu-rep > gpurun_out/r02_ncu_coupling_bwd_merged_kernel_summary.txt; rm -f gpurun_out/r02_prof_coupling_bwd_merged_kernel.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'decoder_eval_tc_kernel|pairwise_cd_fused_kernel|pairwise_emd|cd_scores_kernel' -c 6 -f -o gpurun_out/r02_prof_eval_kernels \
  python tools/eval_kernels_probe.py > gpurun_out/r02_ncu_eval_kernels.log 2>&1; tail -2 gpurun_out/r02_ncu_eval_kernels.log
python tools/ncu_summary.py gpurun_out/r02_prof_eval_kernels.ncu-rep > gpurun_out/r02_ncu_eval_kernels_summary.txt; rm -f gpurun_out/r02_prof_eval_kernels.ncu-rep
# (the .ncu-rep files are ~17 MB each and gpurun_out/ travels back only up to 64 MiB: the text summaries are what is kept)
du -sh gpurun_out
