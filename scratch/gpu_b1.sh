python bench.py --steps 2 --warmup 1 --diffusion-steps 200 --no-cpu-baseline > gpurun_out/b_dmt.json 2> gpurun_out/b_dmt.err; tail -2 gpurun_out/b_dmt.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/b_dmt.json').read().strip().splitlines()[-1])
print('DMT ms/round', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
r=d['roofline']; print({k:r[k] for k in ('kernel','bound','achieved','peak','unit','frac','us_per_launch','share_of_step')}, 'step_us', r['in_stream_step_us'], 'whole-step frac', r['step']['frac'])
for k in r['kernels']: print('  %-34s %-26s n/step %5.1f us %7.1f share %4.1f%% hbm %s tensor %s' % (k['kernel'][:34], k.get('shape',''), k['launches_per_step'], k['us_per_launch'], 100*k['share_of_step'], '%.2f'%k['frac_hbm'] if 'frac_hbm' in k else '-', '%.2f'%k['frac_tensor'] if 'frac_tensor' in k else '-'))
PY
python bench.py --model DMT_WO_EQ --steps 2 --warmup 1 --diffusion-steps 100 --no-cpu-baseline > gpurun_out/b_wo.json 2> gpurun_out/b_wo.err; tail -2 gpurun_out/b_wo.err; python -c "
import json; d=json.loads(open('gpurun_out/b_wo.json').read().strip().splitlines()[-1]); print('WO_EQ ms/round(100 steps)', d['ms_per_step'], 'frac', d['roofline']['step']['frac'], d['roofline']['kernel'], d['roofline']['frac'])"
python bench.py --n-pad 64 --batch 512 --steps 2 --warmup 1 --diffusion-steps 50 --no-cpu-baseline > gpurun_out/b_n64.json 2> gpurun_out/b_n64.err; tail -2 gpurun_out/b_n64.err; python -c "
import json; d=json.loads(open('gpurun_out/b_n64.json').read().strip().splitlines()[-1]); print('N64 ms/round(50 steps)', d['ms_per_step'], 'frac', d['roofline']['step']['frac'], d['roofline']['kernel'], d['roofline']['frac'])"
