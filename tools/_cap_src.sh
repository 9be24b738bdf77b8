timeout 300 ncu --set full --clock-control none --import-source on -k regex:"pl_wgrad_kernel|pl_gemm_kernel" -s 12 -c 6 -f -o gpurun_out/src_pl python tools/encoder_profile.py > gpurun_out/src_pl.log 2>&1
ls -la gpurun_out/src_pl.ncu-rep
