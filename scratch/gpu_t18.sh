timeout 500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for fm in 111 239; do
DS_FUSE_MASK=$fm timeout 300 python bench.py --steps 3 --warmup 1 --diffusion-steps 100 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/round', d['ms_per_step']); print([(k['kernel'][:26], k.get('shape','')[:22], round(k['us_per_launch'],1)) for k in d['roofline']['kernels'] if k['kernel'][:2]=='k_'])"
done
