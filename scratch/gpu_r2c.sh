# round 2, call C: ncu capture of the stand-alone fused coordinate head (same command exited 0 in call B, nothing changed since)
mkdir -p gpurun_out
timeout 300 python scratch/coord_head_time.py > gpurun_out/r2c_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:coord_head -s 3 -c 1 -f -o gpurun_out/r2c_head python scratch/coord_head_time.py > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_plain.log gpurun_out/r2c_ncu.log
