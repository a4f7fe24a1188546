# round 2, call I: BASELINE configs[2] (10 000 molecules, strong scaling) at N = 1, plus wo_eq and n64 single-round lines
mkdir -p gpurun_out
timeout 1500 python bench.py --workload eval10k --steps 1 --warmup 3 --cpu-repeats 0 > gpurun_out/r2i_eval10k_n1.json 2> gpurun_out/r2i_eval10k_n1.err; tail -1 gpurun_out/r2i_eval10k_n1.json | cut -c1-400
timeout 900 python bench.py --workload wo_eq --steps 2 --warmup 3 --cpu-repeats 0 > gpurun_out/r2i_wo_eq_n1.json 2> gpurun_out/r2i_wo_eq_n1.err; tail -1 gpurun_out/r2i_wo_eq_n1.json | cut -c1-300
timeout 900 python bench.py --workload n64 --steps 1 --warmup 3 --cpu-repeats 0 > gpurun_out/r2i_n64_n1.json 2> gpurun_out/r2i_n64_n1.err; tail -1 gpurun_out/r2i_n64_n1.json | cut -c1-300
