#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_emd_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --sweep-clouds 0 > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.log; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_tmp.json"))
print(json.dumps(d["extra"]["emd"]))
PY
