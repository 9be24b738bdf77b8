mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -3 > gpurun_out/final_gpu_suite_1.txt; cat gpurun_out/final_gpu_suite_1.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -3 > gpurun_out/final_gpu_suite_2.txt; cat gpurun_out/final_gpu_suite_2.txt
python __graft_entry__.py --smoke 2>&1 | tail -1
bash tools/gpu_evidence_r02b.sh 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; tail -c 300 gpurun_out/r02_bench_reference_arm.json
