# round 2, call R: source-major attention pass 1: parity (goldens) + A/B
mkdir -p gpurun_out
DS_ATT_P1=1 timeout 900 python -m pytest tests/test_denoiser_gpu.py tests/test_scale_gpu.py -q 2>&1 | tail -3 | tee gpurun_out/r2r_tests.log
DS_ATT_P1=0 timeout 600 python bench.py --diffusion-steps 200 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2r_p0.json; cut -c80-200 gpurun_out/r2r_p0.json
DS_ATT_P1=1 timeout 600 python bench.py --diffusion-steps 200 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2r_p1.json; cut -c80-200 gpurun_out/r2r_p1.json
