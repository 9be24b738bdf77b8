#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu -s -k "speed_vs_eager" 2>&1 | grep "decoder vs\|passed\|failed\|Error" | cut -c1-900
