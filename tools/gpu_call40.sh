#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pointnet_gpu.py tests/test_model_gpu.py -x -q -m gpu -s 2>&1 | grep -v "^Epoch\|^Model saved" | grep "tf32 vs\|passed\|failed\|Error\|assert" | cut -c1-500 | tail
timeout 600 python bench.py --steps 10 --warmup 3 --sweep-clouds 0 > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.log; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_tmp.json"))
print("ms=%.3f"%d["ms_per_step"], json.dumps(d["extra"]["full_model_step"]))
PY
timeout 300 python tools/model_step_probe.py 2>&1 | grep "gpu \|free-running\|encoder train"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pairwise_cd_kernel" -c 1 -f -o gpurun_out/prof_pairwise_cd python tools/chamfer_probe.py 256 > gpurun_out/ncu_cd.log 2>&1; tail -1 gpurun_out/ncu_cd.log
