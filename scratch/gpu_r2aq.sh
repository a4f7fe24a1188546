# round 2: the driver's scaling command at N = 2 (default workload, weak scaling) + the reference arm under torchrun (rank 0 only)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2aq_configs1_n2.json 2> gpurun_out/r2aq_configs1_n2.err; echo rc=$?; tail -1 gpurun_out/r2aq_configs1_n2.json | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --cpu-repeats 0 > gpurun_out/r2aq_reference_n2.json 2> gpurun_out/r2aq_reference_n2.err; echo rc=$?; tail -1 gpurun_out/r2aq_reference_n2.json | cut -c1-300
