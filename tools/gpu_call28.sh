#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu -k "two_tile and bf16x3" -s 2>&1 | grep -E "p2 two|passed|failed|Error" | head -20
