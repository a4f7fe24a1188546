python -m pytest tests -x -q -m gpu 2>&1 | tail -3
run() { python bench.py --steps 3 --warmup 1 --diffusion-steps 100 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/round', d['ms_per_step'])"; }
echo "== OVERLAP=0"; DS_OVERLAP=0 run
for sp in "124,24" "132,16" "116,32" "100,48" "140,8"; do echo "== SPLIT=$sp"; DS_SPLIT=$sp run; done
