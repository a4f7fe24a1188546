#!/usr/bin/env python
"""Benchmark of the DiffSpectra sampling hot path (BASELINE.json: molecules/s, QM9S allspectra, 1000 steps).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--diffusion-steps S]

One bench "step" = one sampling ROUND of the reference eval driver (sampling.py:390-465): a batch of B=1024
synthetic QM9S-shaped molecules (atom counts from the QM9S histogram, allspectra SpecFormer conditioning) taken
through SpecFormer + S=1000 reverse-diffusion steps + post_process.  value = molecules/s over K rounds (whole job,
all ranks); `e2e` = the same through the public API with HOST spectra in pinned memory (H2D inside the timed region)
and the generated molecules copied back to the host.  N > 1: one process per GPU (torchrun), independent shards
(weak scaling), one NCCL all-gather of the packed molecule records per round.

`--impl reference` times the reference's own algorithm on the host CPU cores (the oracle port of the PyTorch path:
/root/reference is a Python repo that is not present on the GPU box) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'molecules/sec (QM9S allspectra, 1000 steps)'
UNIT = 'molecules/s'
VERSION = 'allspectra'


def alg_flops(n, model='DMT'):
    """ALGORITHMIC FLOPs of one molecule x one denoiser step with n atoms: adaLN hoisted
    per molecule, node parts of edge Linears hoisted per atom, directed edges, no credit for redundant reference work
    and no deduction for exploiting symmetry."""
    if model == 'DMT':
        return 2 * (21007360 + 5197568 * n + 1290560 * n * (n - 1))
    # DMT_WO_EQ (models/dmt_wo_eq.py): per molecule: time MLP + 8 x (1536 + 384) adaLN rows + root RBF rows
    m_mol = 17 * 1024 + 1024 * 1024 + 8 * (1536 + 384) * 1024 + 2 * 1024
    # per atom: NodeEmbed, 8 x (qkv, proj, hoisted node2edge halves, FFN, skip), atom head, position head
    m_node = (15 * 512 + 512 * 256) + 8 * (256 * 768 + 256 * 256 + 2 * 256 * 64 + 2 * 256 * 512 + 256 * 64) + \
             (768 * 256 + 256 * 128 + 128 * 6) + (768 * 256 + 256 * 3)
    # per directed edge: root embedding, 8 x (lin_kv_e, q.(k+ek) and alpha (v+ev), FFN, skip), two edge heads
    m_edge = 68 * 64 + 8 * (64 * 512 + 2 * 256 + 2 * 64 * 128 + 64 * 16) + 2 * (192 * 64 + 64 * 32 + 32)
    return 2 * (m_mol + m_node * n + m_edge * n * (n - 1))


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], burst=d['bf16_tflops'], sustained=d.get('bf16_tflops_sustained', d['bf16_tflops']),
                    which='measured (MEASURED_PEAKS.json)')
    return dict(hbm=6650.0, burst=1590.0, sustained=1400.0, which='fallback (B200_PROFILING.md)')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit())
        if not sm:
            return None
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [nm for k, nm in enumerate(names) if any(len(r) >= 7 and r[3 + k].lower().startswith('active') for r in self.rows)]
        mx = max(int(float(r[1])) for r in self.rows if len(r) >= 7)
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': mx, 'reasons': reasons, 'samples': len(sm)}


# ----------------------------------------------------------------------------- the reference arm / CPU baseline
def host_threads():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1: undo that for the CPU arm)."""
    import torch
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def cpu_reference_rate(sample_b=16, sample_steps=4, repeats=1, model='DMT', n_pad=29, all_max=False):
    """Oracle port of the reference PyTorch path (dense restatement, oracle/dense_oracle.py) on the host cores,
    allspectra, on `sample_b` molecules x `sample_steps` denoiser steps (+ one SpecFormer pass per step, as the
    reference recomputes it every call), extrapolated to 1000 steps.  Returns (molecules/s, seconds, description)."""
    import torch
    host_threads()
    from diffspectra_b200.config import get_config
    from diffspectra_b200.model import DMT_B200, DMT_WO_EQ_B200
    from oracle import dense_oracle as O
    from oracle import weights as W
    N_PAD = n_pad
    torch.manual_seed(42)
    sd = (DMT_B200 if model == 'DMT' else DMT_WO_EQ_B200)(get_config(VERSION, device='cpu')).state_dict()
    fwd = O.dmt_forward if model == 'DMT' else O.dmt_wo_eq_forward
    n = W.sample_n_atoms(sample_b, seed=1234, max_n=min(n_pad, 29))
    if all_max:
        n[:] = n_pad
    nm, em = W.make_masks(n, N_PAD)
    ctx = W.synthetic_spectra(sample_b, VERSION, seed=1235)
    table = O.schedule_table(1000)[:: max(1, 1000 // sample_steps)][:sample_steps]
    g = torch.Generator().manual_seed(42)
    z = O.node_noise_from_raw(torch.randn(sample_b, N_PAD, 3, generator=g), torch.randn(sample_b, N_PAD, 6, generator=g), nm)
    ez = O.edge_noise_from_raw(torch.randn(sample_b, 2, N_PAD, N_PAD, generator=g), em)
    raw = [O.draw_step_noise(sample_b, N_PAD, nm, em, generator=g) for _ in range(sample_steps)]

    def denoise(x, ex, nl, cx, cex):          # the reference re-runs SpecFormer inside every call (dmt.py:348-350)
        return fwd(sd, x, nm, em, ex, nl, cx, cex, O.context_embedding(sd, ctx, VERSION))

    best = None
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            O.ancestral_sample(denoise, table, z, ez, nm, em, raw)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    per_step = best / sample_steps
    rate = sample_b / (per_step * 1000.0)
    desc = ('oracle port of the reference PyTorch path (%s), allspectra, B=%d molecules (%s, N_pad=%d) x %d of 1000 '
            'denoiser steps (SpecFormer recomputed each step like the reference), %.2f s/step, extrapolated to 1000 steps'
            % (model, sample_b, 'all n=%d' % n_pad if all_max else 'QM9S histogram', n_pad, sample_steps, per_step))
    return rate, best, desc


def run_reference(args):
    import torch
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = host_threads()
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_rate(8, 1, model=args.model, n_pad=args.n_pad, all_max=args.all_max)
    rates, secs = [], 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r, s, desc = cpu_reference_rate(16, 2, model=args.model, n_pad=args.n_pad, all_max=args.all_max)
        rates.append(r)
        secs += s
        if time.perf_counter() - t0 > 150:
            break
    value = len(rates) * 16 / sum(16 / r for r in rates)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': len(rates),
        'warmup': min(args.warmup, 1), 'ms_per_step': 1000.0 * args.batch / value, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        # same workload keys as the CUDA arm's line; the CPU arm times a bounded sample of it (cpu_baseline.sample)
        'config': {'workload': workload_name(args), 'model': args.model + ' (oracle port of the reference PyTorch path)',
                   'batch_per_gpu': args.batch, 'diffusion_steps': args.diffusion_steps, 'n_pad': args.n_pad,
                   'n_atoms': 'all %d' % args.n_pad if args.all_max else 'QM9S histogram', 'noise': 'torch generator',
                   'sample': desc, 'parallelism': 'host threads: %d' % cores},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': desc},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from diffspectra_b200 import build as B
    from diffspectra_b200.config import get_config
    from diffspectra_b200.distributed import gather_records, pack_records
    from diffspectra_b200.model import DMT_B200, DMT_WO_EQ_B200
    from diffspectra_b200.noise_schedule import NoiseScheduleVP, ancestral_coefficients
    from oracle import weights as W          # synthetic inputs only (atom-count histogram, spectra)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if rank == 0:
        B.build()
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation when NCCL_DEBUG=VERSION is set in the
        # environment; stdout carries exactly ONE JSON line, so the banner goes to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    Bsz, S, N_PAD = args.batch, args.diffusion_steps, args.n_pad
    torch.manual_seed(42)                                   # random-init weights of the reference architecture
    cls = DMT_B200 if args.model == 'DMT' else DMT_WO_EQ_B200
    model = cls(get_config(VERSION, device=str(dev), precision=args.precision)).eval().to(dev)
    eng = model.engine(dev)
    n_atoms = W.sample_n_atoms(Bsz, seed=1234 + rank, max_n=min(N_PAD, 29)).numpy().astype(np.int32)
    if args.all_max:
        n_atoms[:] = N_PAD
    plan = eng.plan(n_atoms, N_PAD)
    spectra_host = [t.pin_memory() for t in W.synthetic_spectra(Bsz, VERSION, seed=1235 + rank)]
    spectra_dev = [t.to(dev) for t in spectra_host]
    ns = NoiseScheduleVP('cosine', continuous_beta_0=0.1, continuous_beta_1=20.)
    coef = ancestral_coefficients(ns, torch.linspace(ns.T, 1e-3, S, device=dev))
    out = (torch.empty(Bsz, N_PAD, 9, device=dev), torch.empty(Bsz, N_PAD, N_PAD, 2, device=dev))
    rec_host = None

    def one_round(r, e2e):
        nonlocal rec_host
        sp = [t.to(dev, non_blocking=True) for t in spectra_host] if e2e else spectra_dev
        ctx_emb = eng.context_embedding(sp)
        eng.sample_loop(plan, ctx_emb, coef, None, None, None, seed=42, gid_base=(r * world + rank) * Bsz,
                        temperature=1.0, use_graph=True, out=out)
        pos, atom, fc, bond = eng.post_process(plan, out[0], out[1])
        rec = pack_records(pos, atom, fc, bond, torch.as_tensor(n_atoms, device=dev))
        if world > 1:
            rec = gather_records(rec)                        # the single collective of the path
        if e2e:
            if rec_host is None or rec_host.shape != rec.shape:
                rec_host = torch.empty(rec.shape, dtype=rec.dtype, pin_memory=True)
            rec_host.copy_(rec, non_blocking=True)
        return rec

    def timed(k, e2e, base):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count()
        e0.record()
        for i in range(k):
            one_round(base + i, e2e)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), eng.launch_count() - l0

    with torch.no_grad():
        for i in range(args.warmup):
            one_round(i, False)
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        ms, launches = timed(args.steps, False, args.warmup)
        clk = clocks.stop() if rank == 0 else None
        one_round(0, True)
        ms_e2e, _ = timed(args.steps, True, args.warmup + args.steps)
        kern = step_kernel_profile(eng, plan, coef, spectra_dev, n_atoms, args) if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    mols = Bsz * world * args.steps
    value = mols / (ms / 1000.0)
    e2e_value = mols / (ms_e2e / 1000.0)
    flops_round = float(sum(alg_flops(int(n), args.model) for n in n_atoms)) * S            # rank 0's shard, denoiser only
    achieved = flops_round * args.steps / (ms / 1000.0) / 1e12                  # TFLOP/s per GPU
    h2d = sum(t.numel() * 4 for t in spectra_host) + n_atoms.nbytes
    d2h = int(rec_host.numel() * rec_host.element_size()) if rec_host is not None else 0
    cpu_rate, cpu_s, cpu_desc = (cpu_reference_rate(16, 2, model=args.model, n_pad=N_PAD, all_max=args.all_max)
                                 if args.cpu_baseline else (None, 0, 'skipped (--no-cpu-baseline)'))
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args), 'model': args.model + '_B200',
                   'batch_per_gpu': Bsz, 'diffusion_steps': S, 'n_pad': N_PAD,
                   'n_atoms': 'all %d' % N_PAD if args.all_max else 'QM9S histogram, mean %.2f' % n_atoms.mean(),
                   'noise': 'device Philox', 'l2': 'inputs_exceed_l2 (per-step working set >> 126 MB)',
                   'parallelism': 'dp%d independent shards + 1 all-gather/round' % world},
        'denoiser_steps_per_s': args.steps * S / (ms / 1000.0),
        'molecule_steps_per_s': value * S,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': d2h,
                'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': int(launches),
        'clocks': clk,
        'roofline': dict(kern['dominant'], step={
            'bound': 'tensor', 'achieved': achieved, 'peak': peaks['sustained'], 'unit': 'TFLOP/s', 'frac': achieved / peaks['sustained'],
            'what': 'whole denoiser step: algorithmic FLOPs per molecule-step (SURVEY.md 8(d); bench.alg_flops) / CUDA-event time '
                    'of the timed rounds; peak = sustained bf16, ' + peaks['which']}, kernels=kern['kernels'],
            in_stream_step_us=kern['step_us']),
        'cpu_baseline': {'value': cpu_rate, 'unit': UNIT, 'cores': host_threads(), 'kind': 'port', 'sample': cpu_desc},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def workload_name(args):
    tag = {29: 'BASELINE.json configs[1]', 64: 'BASELINE.json configs[4] stress shape'}.get(args.n_pad, 'custom shape')
    if args.model != 'DMT':
        tag = 'BASELINE.json configs[3]' if args.n_pad == 29 else tag
    return '%s allspectra sampling, %d steps, batch %d per GPU, N<=%d (%s)' % (args.model, args.diffusion_steps, args.batch,
                                                                              args.n_pad, tag)


def _demangle(name):
    import re
    m = re.search(r'\d+(k_[a-z0-9_]+|gemm_tc_kernel|gemm_simt_kernel|coord_fused_kernel|edge_ffn_kernel)', name)
    base = m.group(1) if m else name[:40]
    t = re.search(r'gemm_tc_kernelILi(\d+)ELi(\d+)ELb(\d)E', name)
    if t:
        base += '<%s,%s,%s>' % (t.group(1), ('STORE', 'LNMOD', 'RESGATE', 'COORD', 'EHEAD')[int(t.group(2))], t.group(3))
    return base


# dram__bytes_read.sum + dram__bytes_write.sum per launch (MB) from the committed `ncu --set full` capture of this
# command at batch 1024 / QM9S histogram (profiles/r1_final3_ncu.md); reported as roofline.traffic for the same workload only
NCU_DRAM_MB = {'k_attention_grp': 212.3, 'k_coord_ln_async': 227.7, 'gemm_tc_kernel<256,COORD,0>': 171.5, 'edge_ffn_kernel': 74.2,
               'gemm_tc_kernel<64,LNMOD,1>': 43.3, 'k_pos_rbf': 2.3}


def step_kernel_profile(eng, plan, coef, spectra_dev, n_atoms, args):
    """Every kernel of a denoiser step timed IN STREAM with CUDA events (ds_profile_begin/end around a non-graph
    ds_sample_loop of 3 steps, warm caches, same inputs as the timed region).  For each kernel: average launch time,
    ALGORITHMIC bytes / FLOPs per launch and the fraction of the HBM / tensor peak they amount to.
    The roofline object reports the kernel with the largest share of the step."""
    import ctypes
    import torch
    from diffspectra_b200 import _lib as L
    peaks = load_peaks()
    lib = L.lib()
    ctx_emb = eng.context_embedding(spectra_dev)
    steps = min(3, coef.shape[0])
    eng.sample_loop(plan, ctx_emb, coef[:steps], None, None, None, seed=1, use_graph=False)      # warm, untimed
    torch.cuda.synchronize()
    L.check(lib.ds_profile_begin(), 'ds_profile_begin')
    eng.sample_loop(plan, ctx_emb, coef[:steps], None, None, None, seed=1, use_graph=False)
    buf = ctypes.create_string_buffer(1 << 16)
    L.check(lib.ds_profile_end(buf, ctypes.c_size_t(len(buf))), 'ds_profile_end')
    Mn, Mp = plan.Mn, plan.Mp
    Md = 2 * Mp
    wo = args.model != 'DMT'
    Me = Md if wo else Mp                       # rows of the edge tensors
    bytes_of = {      # ALGORITHMIC bytes per launch of the graph-side kernels (what must cross HBM once)
        'k_attention_grp': Mp * 1024 + Mp + Mn * (1536 + 256 * 6),          # e0|e1 once per pair, flags, q|k|v, hn fp32+bf16
        'k_wo_attention': Md * 1024 + Mn * (1536 + 512),
        'k_coord_ln': Mp * 512 + Mn * 1024 + Md * 513,                       # gp, ab in; Z + flags out
        'edge_ffn_kernel': Mp * (256 + 256 + 128) + Mn * 256,               # e in, e out, bf16 copy out, hoisted node2edge rows
        'k_edge_update1': Mp * (256 + 256 + 128) + Mn * 256,
        'k_wo_edge_update1': Md * (256 + 256 + 128) + Mn * 512,
        'k_wo_dir_ln1': Md * (256 + 128),
        'k_rbf': Mp * 128 + Mn * 12,
        'k_pos_rbf': Mp * 128 + Md * 4 + Mn * 24,            # X[:, :64] out, directed-edge weights in, positions in/out
        'k_node_ln1': Mn * (1024 + 512), 'k_node_update1': Mn * (2048 + 1024 + 512), 'k_wo_node_update1': Mn * (2048 + 1024 + 512),
        'k_pos_update': Md * 4 + Mn * 24,
        'k_sampler_pairs': Mp * 24, 'k_sampler_nodes': Mn * 108,
    }
    rows, total = [], 0.0
    for line in buf.value.decode().strip().splitlines():
        name, tag, n, us = line.split('\t')
        tag, n, us = int(tag), int(n), float(us)
        total += us
        kname = _demangle(name)
        rec = {'kernel': kname, 'launches_per_step': n / steps, 'us_per_launch': us / n, 'us_per_step': us / steps}
        flops = by = None
        if tag:
            mode, N, K, M = (tag >> 61) & 7, (tag >> 46) & 0x7fff, (tag >> 30) & 0xffff, tag & 0x3fffffff
            rec['shape'] = '[%d,%d]x[%d,%d]' % (M, K, N, K)
            flops = 2.0 * M * N * K
            out_b = {0: M * N * 2, 1: M * N * 2, 2: M * N * (4 + 4 + 2), 3: M * 5, 4: M * 8}[mode]
            by = M * K * 2 + N * K * 2 + out_b
        else:
            for key, v in bytes_of.items():
                if kname.startswith(key):
                    by = v
        t = us / n * 1e-6
        if flops:
            rec['tflops'] = flops / t / 1e12
            rec['frac_tensor'] = rec['tflops'] / peaks['burst']
        if by:
            rec['alg_bytes'] = int(by)
            rec['gbs'] = by / t / 1e9
            rec['frac_hbm'] = rec['gbs'] / peaks['hbm']
        rows.append(rec)
    for r in rows:
        r['share_of_step'] = r['us_per_step'] / (total / steps)
    rows.sort(key=lambda r: -r['us_per_step'])
    top = rows[0]
    tensor_bound = top.get('frac_tensor', 0) > top.get('frac_hbm', 0)
    dominant = {
        'kernel': top['kernel'] + (' ' + top['shape'] if 'shape' in top else ''),
        'bound': 'tensor' if tensor_bound else 'hbm',
        'achieved': top.get('tflops') if tensor_bound else top.get('gbs'),
        'peak': peaks['burst'] if tensor_bound else peaks['hbm'],
        'unit': 'TFLOP/s' if tensor_bound else 'GB/s',
        'frac': top.get('frac_tensor') if tensor_bound else top.get('frac_hbm'),
        'traffic': (NCU_DRAM_MB.get(top['kernel']) * 1e6 if (args.model == 'DMT' and args.batch == 1024 and args.n_pad == 29 and not args.all_max
                                                             and top['kernel'] in NCU_DRAM_MB) else None),
        'us_per_launch': top['us_per_launch'], 'share_of_step': top['share_of_step'],
        'what': 'dominant kernel of a denoiser step by in-stream CUDA-event time; achieved = algorithmic bytes (or FLOPs) per '
                'launch / average launch time; peak = ' + peaks['which'] + ' (burst figures); traffic = ncu dram bytes per launch '
                '(profiles/r1_final3_ncu.md), bytes',
    }
    return {'dominant': dominant, 'kernels': rows[:14], 'step_us': total / steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=1024)
    ap.add_argument('--diffusion-steps', type=int, default=1000)
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--model', default='DMT', choices=['DMT', 'DMT_WO_EQ'])
    ap.add_argument('--n-pad', type=int, default=29, help='padded atoms per molecule (64 = stress shape, implies --all-max)')
    ap.add_argument('--all-max', '--all29', dest='all_max', action='store_true', help='worst case: every molecule has n_pad atoms')
    ap.add_argument('--no-cpu-baseline', dest='cpu_baseline', action='store_false')
    args = ap.parse_args()
    if args.n_pad > 29:
        args.all_max = True            # BASELINE.json configs[4]: synthetic molecules with N atoms each
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
