mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench rc=$?"
cat gpurun_out/bench_fp32.json
tail -5 gpurun_out/bench_fp32.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 520 -c 600 --csv --log-file gpurun_out/launches_fp32.csv python bench.py --steps 2 --warmup 2 --no-extras > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
