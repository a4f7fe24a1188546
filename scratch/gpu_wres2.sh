for w in 1 0; do
for args in "lnmod 40000 128" "lnmod 19073 128" "lnmod 40000 64" "lnmod 40000 256" "store 40000 128"; do
  echo "== WRES=$w $args"; DS_GEMM_WRES=$w CUDA_LAUNCH_BLOCKING=1 timeout 100 python scratch/lnmod_probe.py $args 2>&1 | grep -E "maxerr|illegal|Error" | head -2
done; done
echo "== full gpu tests WRES=1"
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for w in 1 0; do
echo "== bench 20 steps WRES=$w"
DS_GEMM_WRES=$w python bench.py --steps 2 --warmup 2 --diffusion-steps 20 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/round', d['ms_per_step'], 'ms/denoise-step', d['ms_per_step']/20, [ (k['shape'][:12], round(k['ms']*1000,1)) for k in d['roofline']['kernels']])"
done
