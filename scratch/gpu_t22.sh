for ag in 8 4; do
echo "== DS_ATT_G=$ag"
DS_ATT_G=$ag timeout 300 python bench.py --steps 3 --warmup 1 --diffusion-steps 100 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/round', d['ms_per_step']); print([(k['kernel'][:26], k.get('shape','')[:22], round(k['us_per_launch'],1), k['launches_per_step']) for k in d['roofline']['kernels'][:3]])"
done
DS_ATT_G=4 timeout 600 python -m pytest tests/test_denoiser_gpu.py tests/test_scale_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --steps 1 --warmup 3 --diffusion-steps 20 --no-cpu-baseline 2>/dev/null | head -c 300
