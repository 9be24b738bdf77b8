#!/bin/bash
mkdir -p gpurun_out
DPF_LIB_PATH=dpf_nets_b200/_C_stamps/libdpfnets_b200.so timeout 300 python tools/stamp_probe.py 2>&1 | grep -A16 "p2_tc4"
timeout 900 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -3
