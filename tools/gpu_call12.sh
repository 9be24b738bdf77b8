mkdir -p gpurun_out
for b in 16 32 64 128; do
  timeout 600 python bench.py --steps 5 --warmup 3 --precision bf16x3 --no-extras --batch $b > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_b$b.json"))
kc=d["roofline"]["kernel_classes"]
print("B=$b ms=%.2f pts/s=%.3e"%(d["ms_per_step"], d["value"]), {k: round(v["us_per_launch"],1) for k,v in kc.items() if k in ("fwd_stats","fwd_apply","bwd_p1","bwd_p2")})
PY
done
