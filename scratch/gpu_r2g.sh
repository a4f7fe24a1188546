mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_coord_head_gpu.py -q -s -x 2>&1 | tail -8 > gpurun_out/r2g_head.log; cat gpurun_out/r2g_head.log
timeout 300 python scratch/coord_head_time.py > gpurun_out/r2g_time.log 2>&1; tail -8 gpurun_out/r2g_time.log
