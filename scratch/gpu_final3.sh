set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_final3_n1.json 2> gpurun_out/bench_r1_final3_n1.err; tail -2 gpurun_out/bench_r1_final3_n1.err; head -c 400 gpurun_out/bench_r1_final3_n1.json; echo
python bench.py --n-pad 64 --batch 512 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r1_final3_n64.json 2>/dev/null; head -c 300 gpurun_out/bench_r1_final3_n64.json; echo
python bench.py --steps 1 --warmup 1 --diffusion-steps 2 --no-cpu-baseline > gpurun_out/plain3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file gpurun_out/launches_r1_final3.csv python bench.py --steps 1 --warmup 1 --diffusion-steps 2 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_attention_grp|k_coord_ln|gemm_tc_kernel|edge_ffn|k_pos_rbf|k_pos_update|k_sampler|k_node" -s 280 -c 19 -o gpurun_out/prof_r1_final3 python bench.py --steps 1 --warmup 1 --diffusion-steps 2 --no-cpu-baseline > gpurun_out/ncu_full3.log 2>&1; tail -c 200 gpurun_out/ncu_full3.log
