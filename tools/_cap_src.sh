timeout 300 ncu --set full --clock-control none --import-source on -k regex:decoder_eval_tc_kernel -c 1 -f -o gpurun_out/src_eval python tools/eval_kernels_probe.py > gpurun_out/src_eval.log 2>&1
ls -la gpurun_out/src_eval.ncu-rep
python bench.py --steps 5 --warmup 3 --sweep-clouds 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['extra']; print(d['value'], e['sampling'], e['sampling_bf16'])"
