mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/chamfer_probe.py 256 --ref > gpurun_out/chamfer_probe.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pairwise_cd -c 1 -o gpurun_out/prof_pairwise_cd python tools/chamfer_probe.py 64 > gpurun_out/ncu_pairwise.log 2>&1
tail -5 gpurun_out/smoke.log gpurun_out/pytest_gpu.log gpurun_out/chamfer_probe.log
