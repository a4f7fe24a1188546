python -m pytest tests/test_denoiser_gpu.py tests/test_scale_gpu.py -x -q -m gpu 2>&1 | tail -2
DS_FUSE_MASK=15 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_tmp.csv python bench.py --steps 1 --warmup 0 --diffusion-steps 2 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; tail -c 100 gpurun_out/ncu.log
