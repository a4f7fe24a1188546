# round 2, last call: what the driver runs at round end, on the committed tree (built .so as shipped)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2at_bench.json 2>/dev/null; tail -1 gpurun_out/r2at_bench.json | cut -c1-200
