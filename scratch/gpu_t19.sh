timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tests/run_eval_sharded.py 2>&1 | grep -v Warning | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --steps 2 --warmup 3 --diffusion-steps 200 --no-cpu-baseline > gpurun_out/b_n2_200.json 2> gpurun_out/b_n2_200.err; head -c 600 gpurun_out/b_n2_200.json; echo
