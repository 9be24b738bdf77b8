mkdir -p gpurun_out
timeout 200 python tools/merged_probe.py
timeout 600 python -m pytest tests/test_decoder_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|rc=|^FAILED" gpurun_out/pytest_gpu.log | cut -c1-300 | head
timeout 300 python bench.py --steps 10 --warmup 3 --precision bf16x3 --no-extras > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_bf16x3.json"))
kc=d["roofline"]["kernel_classes"]
print("ms=%.2f pts/s=%.3e e2e=%.3e"%(d["ms_per_step"], d["value"], d["e2e"]["value"]), {k: (round(v["us_per_launch"],1) if v["us_per_launch"] else None) for k,v in kc.items()})
PY
