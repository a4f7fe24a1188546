# round 2, call M: full GPU suite + default bench with the fused coordinate head on by default
mkdir -p gpurun_out
rm -f gpurun_out/parity_stats.json
timeout 1500 python -m pytest tests -q -m gpu --maxfail=12 -s 2>&1 | grep -E "PARITY_STATS|passed|failed|FAILED|Error|DS_FUSE" | tail -40 > gpurun_out/r2m_tests.log; tail -12 gpurun_out/r2m_tests.log | cut -c1-400
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; tail -2 gpurun_out/r2m_bench.err; tail -1 gpurun_out/r2m_bench.json | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2m_smoke.log
