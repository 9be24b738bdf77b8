#!/bin/bash
timeout 900 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu -k "full_size_round_trip" 2>&1 | tail -12 | cut -c1-300
