mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_coord_head_gpu.py -q -x 2>&1 | tail -2
timeout 300 python scratch/coord_head_time.py 2>&1 | tail -4 | tee gpurun_out/r2v_time.log
timeout 600 python bench.py --diffusion-steps 200 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2v_bench.json; cut -c80-200 gpurun_out/r2v_bench.json
