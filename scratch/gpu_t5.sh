for pd in 1 0 1 0; do
echo "== bench 100 steps DS_PDL=$pd"
DS_PDL=$pd python bench.py --steps 3 --warmup 1 --diffusion-steps 100 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/round', d['ms_per_step'], 'ms/denoise-step', d['ms_per_step']/100)"
done
