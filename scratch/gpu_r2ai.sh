mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_coord_head_gpu.py -q -x -s 2>&1 | tail -9
timeout 300 python scratch/coord_head_time.py 2>&1 | tail -4 | tee gpurun_out/r2ai_time.log
timeout 900 python -m pytest tests/test_denoiser_gpu.py tests/test_scale_gpu.py tests/test_gemm_gpu.py -q -x 2>&1 | tail -3
timeout 600 python bench.py --diffusion-steps 200 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2ai_bench.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2ai_bench.json'))
print(d['ms_per_step'], [(k['kernel'], round(k['us_per_launch'],1)) for k in d['roofline']['kernels'][:6]])
PY
