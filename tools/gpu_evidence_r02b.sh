#!/bin/bash
# Round-2 evidence pass for the train-mode PointNet kernels and the latent flow layer kernels + the final bench line:
#   gpurun --timeout 1800 -- 'bash tools/gpu_evidence_r02b.sh'
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"pool_forward_kernel|pl_gemm_kernel|pl_wgrad_kernel|pl_bn_bwd_sums" -s 14 -c 12 -f \
  -o gpurun_out/r02_prof_pointnet_train python tools/encoder_profile.py > gpurun_out/r02_ncu_pointnet_train.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_prof_pointnet_train.ncu-rep > gpurun_out/r02_ncu_pointnet_train_kernels_summary.txt; rm -f gpurun_out/r02_prof_pointnet_train.ncu-rep
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"latent_flow" -s 56 -c 4 -f \
  -o gpurun_out/r02_prof_latent_flow python tools/model_profile.py > gpurun_out/r02_ncu_latent_flow.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_prof_latent_flow.ncu-rep > gpurun_out/r02_ncu_latent_flow_kernels_summary.txt; rm -f gpurun_out/r02_prof_latent_flow.ncu-rep
python tools/encoder_profile.py 2>&1 | grep -v Warn > gpurun_out/r02_encoder_train_profile.txt
python tools/model_profile.py 2>&1 | grep -v Warn > gpurun_out/r02_whole_model_profile.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
tail -c 600 gpurun_out/r02_bench_final.json
du -sh gpurun_out
