# round 2, call AB: per-kernel times at batch 256 / 512 / 1024: does the pair chain speed up when e01 (41 / 83 / 166 MB) fits the L2?
mkdir -p gpurun_out
for b in 256 512 1024; do
timeout 600 python bench.py --batch $b --diffusion-steps 100 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2ab_bench_$b.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2ab_bench_$b.json'))
print($b, round(d['ms_per_step'],1), round(d['value'],1), [(k['kernel'][:22], round(k['us_per_launch'],1)) for k in d['roofline']['kernels'][:8]])
PY
done
