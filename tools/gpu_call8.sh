mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for prec in bf16x3; do
  timeout 600 python bench.py --steps 10 --warmup 3 --precision $prec --no-extras > gpurun_out/bench_$prec.json 2> gpurun_out/bench_$prec.err; echo "bench $prec rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$prec.json"))
print("$prec", "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], d["clocks"])
for k,v in d["roofline"]["kernel_classes"].items(): print("   ", k, v)
PY
done
