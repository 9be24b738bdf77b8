mkdir -p gpurun_out
for cfg in kmajor kmajor_bulk mnmajor mnmajor_swapped; do
  timeout 120 python tools/umma_probe.py $cfg >> gpurun_out/umma_probe.log 2>&1; echo "$cfg rc=$?" >> gpurun_out/umma_probe.log
done
cat gpurun_out/umma_probe.log | grep -v Warning | tail -30
