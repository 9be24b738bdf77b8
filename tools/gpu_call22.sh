#!/bin/bash
# PDL A/B on the default and on a side stream
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -4
for cfg in "2=1 " "2=0 " "2=1 --side-stream" "2=0 --side-stream"; do
  set -- $cfg
  timeout 300 python bench.py --steps 20 --warmup 3 --precision bf16x3 --no-extras --lib-option $1 $2 > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "bench rc=$? ($cfg)"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_ab.json"))
print("   ms=%.3f pts/s=%.3e loss=%s"%(d["ms_per_step"], d["value"], d["e2e"]["loss"]))
PY
done
