#!/bin/bash
# PDL A/B: decoder suite, bench with and without programmatic dependent launch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -8
for opt in "2=1" "2=0"; do
  timeout 300 python bench.py --steps 20 --warmup 3 --precision bf16x3 --no-extras --lib-option $opt > gpurun_out/bench_pdl_$opt.json 2> gpurun_out/bench_pdl_$opt.err; echo "bench rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_pdl_$opt.json"))
kc=d["roofline"]["kernel_classes"]
print("$opt ms=%.3f pts/s=%.3e e2e=%.3e"%(d["ms_per_step"], d["value"], d["e2e"]["value"]), {k: (round(v["us_per_launch"],1) if v["us_per_launch"] else None) for k,v in kc.items()})
PY
done
