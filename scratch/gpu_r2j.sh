# round 2, call J: coordinate head v7 (8 build + 8 epilogue warps), whole-loop graph (WHILE node) tests and A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_coord_head_gpu.py -q -s -x 2>&1 | tail -8 > gpurun_out/r2j_head.log; cat gpurun_out/r2j_head.log
timeout 300 python scratch/coord_head_time.py > gpurun_out/r2j_time.log 2>&1; tail -5 gpurun_out/r2j_time.log
timeout 900 python -m pytest tests/test_sampler_gpu.py -q -x 2>&1 | tail -4 | tee gpurun_out/r2j_sampler.log
DS_LOOP_GRAPH=0 timeout 600 python bench.py --diffusion-steps 200 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-260 | tee gpurun_out/r2j_loop0.log
DS_LOOP_GRAPH=1 timeout 600 python bench.py --diffusion-steps 200 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-260 | tee gpurun_out/r2j_loop1.log
DS_FUSE_MASK=255 timeout 600 python bench.py --diffusion-steps 200 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-260 | tee gpurun_out/r2j_fuse255.log
