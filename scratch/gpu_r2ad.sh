# round 2, call AD: `ncu --set full` of the dominant kernel (coord_head_kernel) inside the bench command, final build
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 --diffusion-steps 4 --no-cpu-baseline > /dev/null 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:coord_head_kernel -s 40 -c 1 -o gpurun_out/r2ad_coord_head -f python bench.py --steps 2 --warmup 1 --diffusion-steps 4 --no-cpu-baseline > gpurun_out/r2ad_ncu.log 2>&1; tail -2 gpurun_out/r2ad_ncu.log
