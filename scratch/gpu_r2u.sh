mkdir -p gpurun_out
timeout 300 python scratch/coord_head_time.py 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:coord_head_kernel -s 3 -c 1 -o gpurun_out/r2u_coord_head -f python scratch/coord_head_time.py > gpurun_out/r2u_ncu.log 2>&1; tail -3 gpurun_out/r2u_ncu.log
