for v in a8_c5 a8_c6; do
cp scratch/lib_$v.so diffspectra_b200/libdiffspectra_b200.so; touch diffspectra_b200/libdiffspectra_b200.so
echo "== $v"; python bench.py --steps 3 --warmup 1 --diffusion-steps 100 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/round', d['ms_per_step']); print([(k['kernel'][:22], round(k['us_per_launch'],1)) for k in d['roofline']['kernels'][:4]])"
done
