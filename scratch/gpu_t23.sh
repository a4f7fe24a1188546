for ag in 4 2; do
echo "== DS_ATT_G=$ag"
DS_ATT_G=$ag timeout 300 python bench.py --steps 3 --warmup 1 --diffusion-steps 100 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/round', d['ms_per_step']); print([(k['kernel'][:26], k.get('shape','')[:22], round(k['us_per_launch'],1), k['launches_per_step']) for k in d['roofline']['kernels'][:3]])"
done
DS_ATT_G=2 timeout 600 python -m pytest tests/test_denoiser_gpu.py tests/test_scale_gpu.py -x -q -m gpu 2>&1 | tail -2
