#!/bin/bash
# fused eval decoder: parity vs per-layer form, decoder suite, bench with extras
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu -k "fused_eval" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01_eval.json 2> gpurun_out/bench_r01_eval.log
cat gpurun_out/bench_r01_eval.json
tail -5 gpurun_out/bench_r01_eval.log
