# round 2, multi-GPU: BASELINE configs[2] (10 000 molecules, strong scaling) + the sharded eval driver check under torchrun
# usage: bash scratch/gpu_r2k.sh N
N=$1
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tests/run_eval_sharded.py 2>&1 | grep EVAL_SHARDED | tee gpurun_out/r2k_eval_sharded_n2.log
fi
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N --workload eval10k --steps 1 --warmup 3 --cpu-repeats 0 --no-cpu-baseline > gpurun_out/r2k_eval10k_n$N.json 2> gpurun_out/r2k_eval10k_n$N.err
tail -1 gpurun_out/r2k_eval10k_n$N.json | cut -c1-300
