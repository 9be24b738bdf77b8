#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/model_step_probe.py > gpurun_out/model_step_probe.txt 2>&1
cat gpurun_out/model_step_probe.txt | cut -c1-220
