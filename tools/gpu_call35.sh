#!/bin/bash
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -m gpu -s 2>&1 | grep -v "^Epoch\|^Model saved" | grep "losses eager\|passed\|failed\|Error\|assert" | cut -c1-900 | tail
