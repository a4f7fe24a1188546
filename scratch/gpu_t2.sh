python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for fm in 31 15; do
echo "== bench 20 steps FUSE_MASK=$fm"
DS_FUSE_MASK=$fm python bench.py --steps 2 --warmup 2 --diffusion-steps 20 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/round', d['ms_per_step'], 'ms/denoise-step', d['ms_per_step']/20)"
done
