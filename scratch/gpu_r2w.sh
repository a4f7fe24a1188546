# round 2, call W: one denoiser step under `ncu --set full` with source import (guidance for the attention / edge FFN / lin_edge kernels)
mkdir -p gpurun_out
timeout 300 python bench.py --diffusion-steps 4 --steps 1 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-200
timeout 1200 ncu --set full --clock-control none --import-source on -s 400 -c 80 -o gpurun_out/r2w_step -f python bench.py --diffusion-steps 4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_ncu.log 2>&1; tail -2 gpurun_out/r2w_ncu.log
ls -la gpurun_out/r2w_step.ncu-rep
