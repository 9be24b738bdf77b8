#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/step_profile.py > gpurun_out/step_profile.txt 2>&1
grep -v "^$" gpurun_out/step_profile.txt | cut -c1-80,150-260 | head -60
