#!/bin/bash
# experiment: per-class times and phase stamps with the contended global atomics compiled out (results invalid, timing only)
mkdir -p gpurun_out
export DPF_LIB_PATH=dpf_nets_b200/_C_stamps/libdpfnets_b200.so
timeout 300 python bench.py --steps 20 --warmup 3 --precision bf16x3 --no-extras > gpurun_out/bench_noatom.json 2> gpurun_out/bench_noatom.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_noatom.json"))
kc=d["roofline"]["kernel_classes"]
print("ms=%.3f pts/s=%.3e loss=%s"%(d["ms_per_step"], d["value"], d["e2e"]["loss"]), {k: (round(v["us_per_launch"],1) if v["us_per_launch"] else None) for k,v in kc.items()})
PY
timeout 300 python tools/stamp_probe.py 2>&1 | grep -v "other_tiles\|final_flush" | tee gpurun_out/stamp_probe_noatom.txt
