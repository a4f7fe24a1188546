# round 2, call S: attention softmax over source halves: parity + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_denoiser_gpu.py tests/test_scale_gpu.py tests/test_sampler_gpu.py -q 2>&1 | tail -3 | tee gpurun_out/r2s_tests.log
timeout 600 python bench.py --diffusion-steps 200 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2s_bench.json; cut -c80-200 gpurun_out/r2s_bench.json
