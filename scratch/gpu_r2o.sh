# round 2, call O: LayerNorm-producer node GEMMs: parity + A/B; coordinate head with deeper epilogue unrolling
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_scale_gpu.py -q -s -k fused 2>&1 | grep -E "DS_FUSE|passed|failed|FAILED|timeout" | tee gpurun_out/r2o_fused.log
timeout 600 python -m pytest tests/test_denoiser_gpu.py tests/test_coord_head_gpu.py -q 2>&1 | tail -3 | tee gpurun_out/r2o_den.log
timeout 300 python scratch/coord_head_time.py 2>&1 | tail -1 | tee gpurun_out/r2o_time.log
DS_FUSE_MASK=511 timeout 600 python bench.py --diffusion-steps 200 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2o_511.json; cut -c80-200 gpurun_out/r2o_511.json
DS_FUSE_MASK=1023 timeout 600 python bench.py --diffusion-steps 200 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2o_1023.json; cut -c80-200 gpurun_out/r2o_1023.json
