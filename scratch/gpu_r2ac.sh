# round 2, call AC: launch list + DRAM bytes of the bench command (final build), per profiles/README
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 --diffusion-steps 4 --no-cpu-baseline > gpurun_out/r2ac_bench.json 2>gpurun_out/r2ac_bench.err || exit 1
tail -1 gpurun_out/r2ac_bench.json | cut -c1-160
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1200 -c 400 -o gpurun_out/r2ac_launches -f python bench.py --steps 2 --warmup 1 --diffusion-steps 4 --no-cpu-baseline > gpurun_out/r2ac_ncu.log 2>&1; tail -2 gpurun_out/r2ac_ncu.log
ls -la gpurun_out/r2ac_launches.ncu-rep
