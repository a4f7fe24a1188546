# round 2, call AE: validation of the state to be judged: full GPU suite, default bench line, smoke, reference arm, eval10k at N = 1
mkdir -p gpurun_out
rm -f gpurun_out/parity_stats.json
timeout 1500 python -m pytest tests -q -m gpu --maxfail=12 -s 2>&1 | grep -E "PARITY_STATS|passed|failed|FAILED|Error" | tail -30 > gpurun_out/r2ae_tests.log; tail -3 gpurun_out/r2ae_tests.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2ae_bench.json 2> gpurun_out/r2ae_bench.err; tail -2 gpurun_out/r2ae_bench.err; tail -1 gpurun_out/r2ae_bench.json | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2ae_reference.json 2> gpurun_out/r2ae_reference.err; tail -1 gpurun_out/r2ae_reference.json | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2ae_smoke.log
timeout 1500 python bench.py --workload eval10k --steps 1 --warmup 3 --cpu-repeats 0 --no-cpu-baseline > gpurun_out/r2ae_eval10k_n1.json 2> gpurun_out/r2ae_eval10k_n1.err; tail -1 gpurun_out/r2ae_eval10k_n1.json | cut -c1-300
timeout 600 python bench.py --workload wo_eq --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2ae_wo_eq.json 2>/dev/null; tail -1 gpurun_out/r2ae_wo_eq.json | cut -c1-200
timeout 600 python bench.py --workload n64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2ae_n64.json 2>/dev/null; tail -1 gpurun_out/r2ae_n64.json | cut -c1-200
