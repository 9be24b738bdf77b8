#!/bin/bash
mkdir -p gpurun_out
DPF_LIB_PATH=dpf_nets_b200/_C_stamps/libdpfnets_b200.so timeout 300 python tools/stamp_probe.py > gpurun_out/stamps.txt 2>&1
cat gpurun_out/stamps.txt
