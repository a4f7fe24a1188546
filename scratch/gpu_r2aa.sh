mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_denoiser_gpu.py tests/test_scale_gpu.py -q -x 2>&1 | tail -2
timeout 600 python bench.py --diffusion-steps 200 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2aa_bench.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2aa_bench.json'))
print(d['ms_per_step'], [(k['kernel'], round(k['us_per_launch'],1)) for k in d['roofline']['kernels'][:6]])
PY
