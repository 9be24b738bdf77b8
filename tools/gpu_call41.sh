#!/bin/bash
DPF_LIB_PATH=dpf_nets_b200/_C_stamps/libdpfnets_b200.so timeout 300 python tools/stamp_probe.py 2>&1 | grep "boundary\|order check"
