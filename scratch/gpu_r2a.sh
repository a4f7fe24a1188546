# round 2, call A: full GPU test suite (incl. the parity statistics), default bench line, smoke
mkdir -p gpurun_out
timeout 1300 python -m pytest tests -q -m gpu --maxfail=12 -s 2>&1 | grep -E "PARITY_STATS|passed|failed|FAILED|Error|error|rmsd|rel " | tail -60 > gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -3 gpurun_out/r2a_bench.err
head -c 600 gpurun_out/r2a_bench.json; echo
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 > gpurun_out/r2a_smoke.log
cat gpurun_out/r2a_smoke.log
tail -5 gpurun_out/r2a_tests.log
