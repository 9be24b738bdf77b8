mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_model_gpu.py tests/test_chamfer_gpu.py -m gpu -q > gpurun_out/pytest_model.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_model.log
grep -E "^E  |passed|failed|rc=|^FAILED" gpurun_out/pytest_model.log | cut -c1-300 | head -30
