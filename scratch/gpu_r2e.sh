# round 2, call E: fused coordinate head v3 (CTA-scope waits, pipelined TMEM loads): unit tests, timing, step bench, ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_coord_head_gpu.py -q -s -x 2>&1 | tail -8 > gpurun_out/r2e_head.log; cat gpurun_out/r2e_head.log
timeout 300 python scratch/coord_head_time.py > gpurun_out/r2e_time.log 2>&1; cat gpurun_out/r2e_time.log
timeout 600 python -m pytest tests/test_scale_gpu.py -q -s -k fused 2>&1 | grep -E "DS_FUSE|passed|failed" | tee gpurun_out/r2e_fused.log
DS_FUSE_MASK=255 timeout 600 python bench.py --diffusion-steps 100 --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2e_bench255.json; cut -c1-200 gpurun_out/r2e_bench255.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:coord_head -s 3 -c 1 -f -o gpurun_out/r2e_head python scratch/coord_head_time.py > gpurun_out/r2e_ncu.log 2>&1; tail -2 gpurun_out/r2e_ncu.log
