#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -m gpu -s 2>&1 | grep -v "^Epoch\|^Model saved" | tail -25
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_full2.json 2> gpurun_out/bench_r01_full2.log; echo "bench rc=$?"
tail -3 gpurun_out/bench_r01_full2.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r01_full2.json"))
kc=d["roofline"]["kernel_classes"]
print("   ms=%.3f pts/s=%.3e e2e=%.3e loss=%s"%(d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["loss"]), {k: (round(v["us_per_launch"],1) if v["us_per_launch"] else None) for k,v in kc.items()})
print(json.dumps(d["extra"]["full_model_step"]))
print(json.dumps(d["extra"]["eval_sweep"]))
print(json.dumps(d["extra"]["chamfer"]["value"]), json.dumps(d["extra"]["sampling"]))
PY
