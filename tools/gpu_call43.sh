#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --sweep-clouds 0 > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.log; echo "bench rc=$?"; tail -3 gpurun_out/bench_tmp.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_tmp.json"))
print(json.dumps(d["extra"]["emd"])); print(d["roofline"]["traffic"], d["ms_per_step"])
PY
