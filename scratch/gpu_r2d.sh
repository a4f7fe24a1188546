# round 2, call D: fused coordinate head v2 (two threads per edge, pipelined passes): unit tests, timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_coord_head_gpu.py -q -s -x 2>&1 | tail -12 > gpurun_out/r2d_head.log; cat gpurun_out/r2d_head.log
timeout 300 python scratch/coord_head_time.py 2>&1 | tail -3 | tee gpurun_out/r2d_time.log
timeout 600 python -m pytest tests/test_scale_gpu.py -q -s -k fused 2>&1 | tail -4 | tee gpurun_out/r2d_fused.log
DS_FUSE_MASK=255 timeout 600 python bench.py --diffusion-steps 100 --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2d_bench255.json; cut -c1-300 gpurun_out/r2d_bench255.json
