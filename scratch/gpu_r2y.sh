# round 2, call Y: tanh epilogue of lin_edge with half of the column pairs on the FMA pipe (DS_TANH_MIX) A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -x -k "tanh or epilogue or many_tiles" 2>&1 | tail -3
for m in 0 1; do
DS_TANH_MIX=$m timeout 600 python bench.py --diffusion-steps 200 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2y_bench_$m.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2y_bench_$m.json'))
print('mix=$m', d['ms_per_step'], [(k['kernel'], round(k['us_per_launch'],1)) for k in d['roofline']['kernels'][:5]])
PY
done
timeout 900 python -m pytest tests/test_parity_stats_gpu.py tests/test_denoiser_gpu.py -q -x 2>&1 | tail -3
