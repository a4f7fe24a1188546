set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -4
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; tail -c 3000 gpurun_out/bench_r1_n1.json; tail -3 gpurun_out/bench_r1_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err; cat gpurun_out/bench_r1_ref.json; tail -3 gpurun_out/bench_r1_ref.err
python bench.py --steps 1 --warmup 1 --diffusion-steps 2 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 1 --warmup 1 --diffusion-steps 2 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; tail -c 200 gpurun_out/ncu.log
