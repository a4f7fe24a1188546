for sp in 1 2 3 4; do
echo "== DS_L2_SPLIT=$sp"
DS_L2_SPLIT=$sp timeout 300 python bench.py --steps 3 --warmup 1 --diffusion-steps 100 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/round', d['ms_per_step']); print([(k['kernel'][:26], k.get('shape','')[:22], round(k['us_per_launch'],1), k['launches_per_step']) for k in d['roofline']['kernels'][:7]])"
done
DS_L2_SPLIT=2 timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
