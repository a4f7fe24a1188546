for fv in 0 1; do
echo "== DS_FFN=$fv"
DS_FFN=$fv timeout 300 python bench.py --steps 3 --warmup 1 --diffusion-steps 100 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/round', d['ms_per_step']); print([(k['kernel'][:26], round(k['us_per_launch'],1)) for k in d['roofline']['kernels'] if 'ffn' in k['kernel']])"
DS_FFN=$fv timeout 600 python -m pytest tests/test_denoiser_gpu.py tests/test_scale_gpu.py -x -q -m gpu 2>&1 | tail -1
done
