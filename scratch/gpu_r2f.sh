mkdir -p gpurun_out
timeout 300 python scratch/coord_head_time.py > gpurun_out/r2f_time.log 2>&1; cat gpurun_out/r2f_time.log
