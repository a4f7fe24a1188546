for b in 1024 2048 4096 8192; do
python bench.py --batch $b --steps 2 --warmup 1 --diffusion-steps 50 --no-cpu-baseline 2>&1 | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); b=d['config']['batch_per_gpu']; print('batch', b, 'ms/round(50 steps)', round(d['ms_per_step'],1), 'in-stream step us', round(d['roofline']['in_stream_step_us']), 'us per molecule-step', round(d['roofline']['in_stream_step_us']/b,3), 'frac', round(d['roofline']['step']['frac'],3))"
done
