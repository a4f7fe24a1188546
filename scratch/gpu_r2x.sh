# round 2, call X: `ncu --set full` with source of ONE attention and ONE edge-FFN launch inside a denoiser step
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_attention_grp|edge_ffn_kernel" -s 40 -c 2 -o gpurun_out/r2x_att_ffn -f python bench.py --diffusion-steps 4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2x_ncu.log 2>&1; tail -2 gpurun_out/r2x_ncu.log
ls -la gpurun_out/r2x_att_ffn.ncu-rep
