mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_coord_head_gpu.py -q -s -x 2>&1 | tail -8 > gpurun_out/r2l_head.log; cat gpurun_out/r2l_head.log
timeout 300 python scratch/coord_head_time.py > gpurun_out/r2l_time.log 2>&1; tail -5 gpurun_out/r2l_time.log
DS_FUSE_MASK=255 timeout 600 python bench.py --diffusion-steps 200 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-260 | tee gpurun_out/r2l_fuse255.log
