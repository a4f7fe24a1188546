"""profiles/r2_ncu_coord_head.md from the `ncu --set full --import-source on` capture of coord_head_kernel (scratch/gpu_r2ad.sh):

    ncu -i gpurun_out/r2ad_coord_head.ncu-rep --page raw --csv    > gpurun_out/r2ad_raw.csv
    ncu -i gpurun_out/r2ad_coord_head.ncu-rep --page source --csv > gpurun_out/r2ad_src.csv
    python scratch/ncu_coord_head_md.py > profiles/r2_ncu_coord_head.md
"""
import csv
rows = list(csv.reader(open('gpurun_out/r2ad_raw.csv')))
h, u, v = rows[0], rows[1], rows[2]
def g(k): return v[h.index(k)] + ' ' + u[h.index(k)]
print('# Round 2: `ncu --set full` of the dominant kernel, final build (coordinate head v10)\n')
print('Command (`scratch/gpu_r2ad.sh`, after the same bench command had exited 0 without ncu):\n')
print('    ncu --set full --clock-control none --import-source on -k regex:coord_head_kernel -s 40 -c 1 \\\n        python bench.py --steps 2 --warmup 1 --diffusion-steps 4 --no-cpu-baseline\n')
print('`coord_head_kernel` (coord_head_tc.cu; grid 148 = 74 CTA pairs, 640 threads, 229.6 KB shared memory, 512 TMEM columns per CTA) at the')
print('configs[1] shape (Mp = 162 305 pairs, Md = 324 610 directed edges).  A number under ncu is not a bench value: the in-stream')
print('CUDA-event time of the same launch is in `profiles/r2_bench_final_configs1.json` (roofline.kernels).\n')
print('| metric | value |\n|---|---|')
for k in ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
          'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
          'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum',
          'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active']:
    if k in h: print('| `%s` | %s |' % (k, g(k)))
print('\nWarp stall reasons (warps per issue-active cycle): ' + '; '.join('%s %s' % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v[i][:5]) for i, k in enumerate(h) if 'issue_stalled' in k and 'per_issue_active' in k and float(v[i]) >= 0.15))
print('''
Reading:

* DRAM traffic per launch against 62.0 MB algorithmic (X 41.5 MB + per-atom table 19 MB, L2-resident after the first touch, +
  1.3 MB of results): the G and Z round trips of the three kernels this one replaced (~580 MB) are gone.
* No pipe is saturated; the most loaded unit is the L1 data pipe (the epilogue's fp32 table, the operand tile Z, the per-atom and
  modulate rows).  The kernel is a two-stage software pipeline (build warps: TMEM -> LayerNorm -> Z; epilogue warps: TMEM ->
  SiLU -> 3 dot products) with 2 + 2 warps per scheduler; the register file (640 threads x 96) caps the warp count.
* History of the capture: v7 (stand-alone, 118.8 us under ncu) had long scoreboard 4.4, membar 0.68 (cluster-scope release in front
  of every tmem_free arrive), barrier 2.5 and 2.6 M + 2.3 M L1 sectors of local memory (spills); v9 3.8 / 0.16 / 2.0 and
  0.57 M + 0.18 M, with 30 % of the build warps' samples waiting for G and 41 % of the epilogue warps' waiting for the accumulator
  (shared TMEM stages); v10 (this capture) has G and the accumulator halves in separate TMEM columns (DESIGN.md, round-2 table).
''')
rows = list(csv.reader(open('gpurun_out/r2ad_src.csv')))[2:]
tot = sum(int(r[2]) for r in rows)
print('Stall samples by SASS region (%d samples; region = 40 consecutive instructions, opcode counts of the region):\n' % tot)
print('| first instr | samples | max executed | opcodes |\n|---|---|---|---|')
W = 40
for b in range(0, len(rows), W):
    seg = rows[b:b + W]; s = sum(int(r[2]) for r in seg); ex = max(int(r[5]) for r in seg); ops = {}
    for r in seg:
        op = [o for o in r[1].strip().split() if not o.startswith('@')]
        op = op[0].split('.')[0] if op else ''
        if op in ('MUFU', 'UTCHMMA', 'UTMALDG', 'HFMA2', 'LDS', 'STS', 'LDTM', 'SYNCS', 'BAR', 'LDG', 'UTCBAR', 'F2FP', 'FFMA2', 'FADD2', 'MEMBAR', 'ERRBAR', 'FENCE', 'STG', 'LDL', 'STL'): ops[op] = ops.get(op, 0) + 1
    if s >= 60: print('| %d | %d | %d | %s |' % (b, s, ex, ' '.join('%s %d' % (k, c) for k, c in ops.items())))
