#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chamfer_gpu.py -x -q -m gpu -s -k "speed_vs_reference" 2>&1 | grep "chamfer vs\|passed\|failed\|Error" | cut -c1-900
