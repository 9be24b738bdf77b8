#!/bin/bash
# Quick A/B loop on a B200 box: decoder parity tests + a short headline bench with the per-kernel-class table.
#   gpurun --timeout 900 -- 'bash tools/gpu_quick.sh [--lib-option K=V ...]'
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras "$@" > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_ab.json"))
kc = d["roofline"]["kernel_classes"]
print("ms=%.3f pts/s=%.3e e2e=%.3e loss=%s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["loss"]),
      {k: (round(v["us_per_launch"], 1) if v["us_per_launch"] else None) for k, v in kc.items()})
PY
