mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_emd_gpu.py -m gpu -q > gpurun_out/pytest_emd.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_emd.log
grep -E "^E  |passed|failed|rc=" gpurun_out/pytest_emd.log | cut -c1-250 | head -30
