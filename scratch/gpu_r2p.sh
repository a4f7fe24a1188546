# round 2, call P: once-per-step kernels (root nodes, node sampler with three warps per molecule, time MLP once per step)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sampler_gpu.py tests/test_denoiser_gpu.py tests/test_scale_gpu.py tests/test_wo_eq_gpu.py -q 2>&1 | tail -4 | tee gpurun_out/r2p_tests.log
timeout 600 python bench.py --diffusion-steps 200 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2p_bench.json; cut -c80-200 gpurun_out/r2p_bench.json
