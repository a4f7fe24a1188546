set -x
timeout 900 python -m pytest tests -x -q -m gpu -s 2>&1 | grep -E "passed|failed|x_mean rel|rmsd per molecule|wo_eq|error" | tail -30
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_final2_n1.json 2> gpurun_out/bench_r1_final2_n1.err; tail -2 gpurun_out/bench_r1_final2_n1.err; head -c 400 gpurun_out/bench_r1_final2_n1.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_final2_ref.json 2>/dev/null; head -c 300 gpurun_out/bench_r1_final2_ref.json; echo
python bench.py --model DMT_WO_EQ --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r1_final2_wo_eq.json 2>/dev/null; head -c 300 gpurun_out/bench_r1_final2_wo_eq.json; echo
python bench.py --n-pad 64 --batch 512 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r1_final2_n64.json 2>/dev/null; head -c 300 gpurun_out/bench_r1_final2_n64.json; echo
python bench.py --steps 1 --warmup 1 --diffusion-steps 2 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file gpurun_out/launches_r1_final2.csv python bench.py --steps 1 --warmup 1 --diffusion-steps 2 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_attention_grp|k_coord_ln|gemm_tc_kernel|edge_ffn|k_pos_rbf|k_pos_update|k_sampler|k_node" -s 280 -c 24 -o gpurun_out/prof_r1_final2 python bench.py --steps 1 --warmup 1 --diffusion-steps 2 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1; tail -c 200 gpurun_out/ncu_full2.log
