mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -k "bf16 or umma" > gpurun_out/pytest_bf16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_bf16.log
tail -40 gpurun_out/pytest_bf16.log
