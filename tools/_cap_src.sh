timeout 300 ncu --set full --clock-control none --import-source on -k regex:"latent_flow" -s 30 -c 2 -f -o gpurun_out/src_lf python tools/model_profile.py > gpurun_out/src_lf.log 2>&1
ls -la gpurun_out/src_lf.ncu-rep
