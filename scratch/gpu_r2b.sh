# round 2, call B: CTA-pair probe, fused coordinate head unit tests, stand-alone timing, fp32-mode step time
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_coord_head_gpu.py -q -s -k probe 2>&1 | tail -15 > gpurun_out/r2b_probe.log; cat gpurun_out/r2b_probe.log
timeout 600 python -m pytest tests/test_coord_head_gpu.py -q -s -x -k coord_head 2>&1 | tail -30 > gpurun_out/r2b_head.log; cat gpurun_out/r2b_head.log
timeout 300 python scratch/coord_head_time.py 2>&1 | tail -3 | tee gpurun_out/r2b_time.log
timeout 600 python -m pytest tests/test_scale_gpu.py -q -s -k fused 2>&1 | tail -8 | tee gpurun_out/r2b_fused.log
timeout 600 python bench.py --precision fp32 --diffusion-steps 20 --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-400 | tee gpurun_out/r2b_fp32.log
