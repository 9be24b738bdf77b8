#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r01_n2.json 2> gpurun_out/bench_r01_n2.log; echo "bench n2 rc=$?"
tail -5 gpurun_out/bench_r01_n2.log | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r01_n2.json"))
print("n_gpus", d["n_gpus"], "ms=%.3f pts/s=%.3e e2e=%.3e"%(d["ms_per_step"], d["value"], d["e2e"]["value"]))
print(json.dumps(d["extra"]["eval_sweep"]))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | cut -c1-300
