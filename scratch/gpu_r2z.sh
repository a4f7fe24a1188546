mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -x 2>&1 | tail -2
timeout 300 python scratch/lin_edge_time.py 2>&1 | tail -5
timeout 600 python bench.py --diffusion-steps 200 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2z_bench.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2z_bench.json'))
print(d['ms_per_step'], [(k['kernel'], round(k['us_per_launch'],1)) for k in d['roofline']['kernels'][:14]])
PY
