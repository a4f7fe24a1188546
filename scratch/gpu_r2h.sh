mkdir -p gpurun_out
timeout 300 python scratch/coord_head_time.py > gpurun_out/r2h_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:coord_head -s 3 -c 1 -f -o gpurun_out/r2h_head python scratch/coord_head_time.py > gpurun_out/r2h_ncu.log 2>&1; tail -2 gpurun_out/r2h_ncu.log
