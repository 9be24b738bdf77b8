mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -2
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 200 gpurun_out/r02_bench_final.json
python tools/model_profile.py 2>&1 | grep -v Warn > gpurun_out/r02_whole_model_profile.txt
python tools/encoder_profile.py 2>&1 | grep -v Warn > gpurun_out/r02_encoder_train_profile.txt
python tools/latent_flow_stamps.py > gpurun_out/r02_latent_flow_phase_stamps.txt 2>&1
