#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -m gpu -s -k "graphed" 2>&1 | grep -v "^Epoch\|^Model saved" | tail -25 | cut -c1-400

