# round 2, call N: skip projection folded into the fused edge stream: parity + A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_scale_gpu.py tests/test_denoiser_gpu.py -q -s 2>&1 | grep -E "DS_FUSE|passed|failed|FAILED" | tee gpurun_out/r2n_tests.log
DS_FUSE_MASK=255 timeout 600 python bench.py --diffusion-steps 200 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c80-200 | tee gpurun_out/r2n_255.log
DS_FUSE_MASK=511 timeout 600 python bench.py --diffusion-steps 200 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c80-200 | tee gpurun_out/r2n_511.log
