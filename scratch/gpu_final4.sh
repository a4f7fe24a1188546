timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_final4_n1.json 2> gpurun_out/bench_r1_final4_n1.err; tail -2 gpurun_out/bench_r1_final4_n1.err; head -c 300 gpurun_out/bench_r1_final4_n1.json; echo
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
