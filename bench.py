#!/usr/bin/env python
"""Benchmark of the DiffSpectra sampling hot path (BASELINE.json: molecules/s, QM9S allspectra, 1000 steps).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload configs1|eval10k|wo_eq|n64]

One bench "step" = one sampling ROUND of the reference eval driver (sampling.py:390-465): SpecFormer + S=1000
reverse-diffusion steps + molecule records.  Workloads (BASELINE.json `configs`):
  configs1 (default)  configs[1]: DMT allspectra, batch 1024 per GPU, QM9S atom-count histogram; weak scaling
  eval10k             configs[2]: 10 000 samples x 1000 steps, STRONG scaling: every rank samples ceil(10000/G) molecules
                      in ONE round (1250 per GPU at 8), one all-gather of the records inside the timed region
  wo_eq               configs[3]: the DMT_WO_EQ ablation, batch 1024 per GPU
  n64                 configs[4]: stress shape, 64 atoms per molecule, batch 512 per GPU
`value` = molecules/s with the spectra already resident in HBM (Engine calls: SpecFormer -> loop -> records -> gather);
`e2e` = the same through the PUBLIC API — `get_cond_sampling_eval_fn(config, ...)(model)` (sampling.py:353-468) on an
in-memory data set with HOST spectra: staging, H2D, plan build, weight fingerprint, SpecFormer, the loop, records, the
all-gather, ONE D2H copy and the conversion to the reference's per-molecule tuples all inside the timed region.

`--impl reference` times the reference's own algorithm on the host CPU cores (the oracle port of the PyTorch path:
/root/reference is a Python repo that is not present on the GPU box) on a bounded sample of the same workload.
The CUDA arm imports nothing from oracle/ except for the `cpu_baseline` leg.
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


def cpu_configs0(repeats=3):
    """BASELINE.json configs[0] AS WRITTEN: reference path (oracle port), IR-only SpecFormer, B = 16
    molecules (QM9S histogram, molecule 0 with 29 atoms), N_pad = 29, 50-step ancestral sampling, fp32, timed in full
    (no extrapolation); `repeats` runs after one warm-up denoiser call.  Returns dict(min/median/max seconds, rates)."""
    import torch
    host_threads()
    from diffspectra_b200.config import get_config
    from diffspectra_b200.model import DMT_B200
    from oracle import dense_oracle as O
    from oracle import weights as W
    B, N, steps, version = 16, 29, 50, 'ir'
    torch.manual_seed(42)
    sd = DMT_B200(get_config(version, device='cpu')).state_dict()
    n = W.sample_n_atoms(B, seed=1234)
    nm, em = W.make_masks(n, N)
    ctx = W.synthetic_spectra(B, version, seed=1235)
    table = O.schedule_table(steps)
    g = torch.Generator().manual_seed(42)
    z = O.node_noise_from_raw(torch.randn(B, N, 3, generator=g), torch.randn(B, N, 6, generator=g), nm)
    ez = O.edge_noise_from_raw(torch.randn(B, 2, N, N, generator=g), em)
    raw = [O.draw_step_noise(B, N, nm, em, generator=g) for _ in range(steps)]

    def denoise(x, ex, nl, cx, cex):          # the reference re-runs SpecFormer inside every call (dmt.py:348-350)
        return O.dmt_forward(sd, x, nm, em, ex, nl, cx, cex, O.context_embedding(sd, ctx, version))

    secs = []
    with torch.no_grad():
        denoise(z, ez, torch.full((B,), -9.5), None, None)        # warm-up call
        for _ in range(repeats):
            t0 = time.perf_counter()
            O.ancestral_sample(denoise, table, z, ez, nm, em, raw)
            secs.append(time.perf_counter() - t0)
    secs.sort()
    return {'workload': 'BASELINE configs[0]: DMT + IR-only SpecFormer, B=16 (QM9S histogram), N_pad=29, 50-step ancestral '
                        'sampling, fp32, timed in full', 'repeats': repeats, 'seconds_min': secs[0],
            'seconds_median': secs[len(secs) // 2], 'seconds_max': secs[-1], 'molecules_per_s_best': B / secs[0],
            'denoiser_steps_per_s_best': steps / secs[0], 'spread': (secs[-1] - secs[0]) / secs[0]}


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


WORKLOADS = {
    # name: (model, per-GPU batch (None: strong), n_pad, all_max, total molecules (strong) or None, BASELINE tag)
    'configs1': dict(model='DMT', batch=1024, n_pad=29, all_max=False, total=None, tag='BASELINE.json configs[1]'),
    'eval10k': dict(model='DMT', batch=None, n_pad=29, all_max=False, total=10000, tag='BASELINE.json configs[2]'),
    'wo_eq': dict(model='DMT_WO_EQ', batch=1024, n_pad=29, all_max=False, total=None, tag='BASELINE.json configs[3]'),
    'n64': dict(model='DMT', batch=512, n_pad=64, all_max=True, total=None, tag='BASELINE.json configs[4] stress shape'),
}


def resolve_workload(args, world):
    w = dict(WORKLOADS[args.workload])
    if args.model is not None:
        w['model'] = args.model
    if args.n_pad is not None:
        w['n_pad'] = args.n_pad
        w['all_max'] = w['all_max'] or args.n_pad > 29
    if args.all_max:
        w['all_max'] = True
    if args.batch is not None:
        w['batch'] = args.batch
        w['total'] = None
    if w['total'] is not None:                     # strong scaling: one round of ceil(total / G) molecules per rank
        w['batch'] = -(-w['total'] // world)
    w['strong'] = w['total'] is not None
    if args.model is not None or args.n_pad is not None or args.batch is not None or args.all_max:
        w['tag'] = 'custom (derived from %s)' % w['tag']
    return w


def workload_name(args, w=None):
    w = w or resolve_workload(args, max(1, args.gpus))
    if w['strong']:
        return '%s allspectra sampling, %d steps, %d molecules in total sharded over the GPUs (%d per GPU, one round), N<=%d (%s)' % (
            w['model'], args.diffusion_steps, w['total'], w['batch'], w['n_pad'], w['tag'])
    return '%s allspectra sampling, %d steps, batch %d per GPU, N<=%d (%s)' % (w['model'], args.diffusion_steps, w['batch'],
                                                                              w['n_pad'], w['tag'])


def run_reference(args):
    import torch
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    w = resolve_workload(args, max(1, args.gpus))
    cores = host_threads()
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_rate(8, 1, model=w['model'], n_pad=w['n_pad'], all_max=w['all_max'])
    rates, secs = [], 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r, sec, desc = cpu_reference_rate(16, 2, model=w['model'], n_pad=w['n_pad'], all_max=w['all_max'])
        rates.append(r)
        secs += sec
        if time.perf_counter() - t0 > 150:
            break
    value = len(rates) * 16 / sum(16 / r for r in rates)
    c0 = cpu_configs0(repeats=args.cpu_repeats) if (w['model'] == 'DMT' and w['n_pad'] == 29 and args.cpu_repeats > 0) else None
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': len(rates),
        'warmup': min(args.warmup, 1), 'ms_per_step': 1000.0 * w['batch'] / value, 'higher_is_better': True,
        'scaling': 'strong' if w['strong'] else 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        # same workload keys as the CUDA arm's line; the CPU arm times a bounded sample of it (cpu_baseline.sample)
        'config': {'workload': workload_name(args, w), 'model': w['model'] + ' (oracle port of the reference PyTorch path)',
                   'batch_per_gpu': w['batch'], 'diffusion_steps': args.diffusion_steps, 'n_pad': w['n_pad'],
                   'n_atoms': 'all %d' % w['n_pad'] if w['all_max'] else 'QM9S histogram', 'noise': 'torch generator',
                   'sample': desc, 'parallelism': 'host threads: %d' % cores},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': desc,
                         'value_min': min(rates), 'value_max': max(rates), 'configs0': c0},
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
    from diffspectra_b200 import evaluate as E
    from diffspectra_b200 import sampling as SMP
    from diffspectra_b200 import synthetic as SY
    from diffspectra_b200.config import get_config
    from diffspectra_b200.distributed import gather_records, record_bytes
    from diffspectra_b200.model import DMT_B200, DMT_WO_EQ_B200
    from diffspectra_b200.noise_schedule import NoiseScheduleVP, ancestral_coefficients

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

    w = resolve_workload(args, world)
    Bsz, S, N_PAD = w['batch'], args.diffusion_steps, w['n_pad']
    torch.manual_seed(42)                                   # random-init weights of the reference architecture
    cls = DMT_B200 if w['model'] == 'DMT' else DMT_WO_EQ_B200
    config = get_config(VERSION, device=str(dev), precision=args.precision)
    config.sampling.steps = S
    config.data.max_node = N_PAD
    model = cls(config).eval().to(dev)
    eng = model.engine(dev)
    # the data set of the whole job (all ranks build the same one): rank r's shard = items [r*Bsz, (r+1)*Bsz) of the
    # driver's permutation; the device-resident leg uses the same atom counts / spectra without the permutation
    n_items = Bsz * world
    n_all = SY.sample_n_atoms(n_items, seed=1234, max_n=min(N_PAD, 29))
    if w['all_max']:
        n_all[:] = N_PAD
    if w['strong'] and w['total'] is not None:
        n_items = w['total']
        n_all = n_all[:n_items]
    ds = SY.SyntheticQM9S(n_items, VERSION, seed=1235, n_atoms=n_all, pin=True)
    lo = min(rank * Bsz, n_items)
    hi = min(lo + Bsz, n_items)
    n_atoms = n_all[lo:hi].numpy().astype(np.int32)
    B_loc = int(hi - lo)
    plan = eng.plan(n_atoms, N_PAD)
    spectra_dev = [getattr(ds._data, k)[lo:hi].to(dev) for k in ds.keys]
    ns = NoiseScheduleVP('cosine', continuous_beta_0=0.1, continuous_beta_1=20.)
    coef = ancestral_coefficients(ns, torch.linspace(ns.T, 1e-3, S, device=dev))
    out = (torch.empty(B_loc, N_PAD, 9, device=dev), torch.empty(B_loc, N_PAD, N_PAD, 2, device=dev))
    pad = None
    if B_loc < Bsz:                                          # short last shard of the strong-scaling split
        pad = torch.zeros(Bsz - B_loc, record_bytes(N_PAD), dtype=torch.uint8, device=dev)

    def resident_round(r):
        """Inputs already in HBM: SpecFormer -> 1000-step loop -> record kernel -> (all-gather)."""
        ctx_emb = eng.context_embedding(spectra_dev)
        eng.sample_loop(plan, ctx_emb, coef, None, None, None, seed=42, gid_base=(r * world + rank) * Bsz,
                        temperature=1.0, use_graph=True, out=out)
        rec = eng.molecule_records(plan, out[0], out[1], N_PAD)
        if pad is not None:
            rec = torch.cat([rec, pad])
        if world > 1:
            rec = gather_records(rec)                        # the single collective of the path
        return rec

    # the public API: the reference's eval-driver factory (sampling.py:353) -> sampling_fn(model)
    if world > 1:
        api_fn = E.get_cond_sampling_eval_fn(config, ns, Bsz, n_items, None, ds, noise='philox', seed=42)
    else:
        api_fn = SMP.get_cond_sampling_eval_fn(config, ns, Bsz, n_items, None, ds, noise='philox', seed=42)
    n_out = [0]

    def api_round(r):
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):        # the driver prints progress like the reference does (sampling.py:463);
            mols, tpos, _ = api_fn(model)                   # stdout carries exactly ONE JSON line
        n_out[0] = len(mols)
        return mols

    def timed(k, fn, base):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count()
        t0 = time.perf_counter()
        e0.record()
        for i in range(k):
            fn(base + i)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1000.0
        ms = torch.tensor([max(e0.elapsed_time(e1), 0.0), wall], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms[0].item(), ms[1].item(), eng.launch_count() - l0

    with torch.no_grad():
        for i in range(args.warmup):
            resident_round(i)
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        ms, _, launches = timed(args.steps, resident_round, args.warmup)
        clk = clocks.stop() if rank == 0 else None
        api_round(0)                                        # warm the API path (pinned buffers, caches)
        ms_e2e_dev, ms_e2e_wall, _ = timed(args.steps, api_round, 1)
        ms_e2e = max(ms_e2e_dev, ms_e2e_wall)               # the host tail (records -> tuples) is part of the call
        kern = step_kernel_profile(eng, plan, coef, spectra_dev, n_atoms, args, w) if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    mols = n_items * args.steps
    value = mols / (ms / 1000.0)
    e2e_value = mols / (ms_e2e / 1000.0)
    assert n_out[0] == n_items, (n_out[0], n_items)
    flops_round = float(sum(alg_flops(int(n), w['model']) for n in n_atoms)) * S            # rank 0's shard, denoiser only
    achieved = flops_round * args.steps / (ms / 1000.0) / 1e12                  # TFLOP/s per GPU
    h2d = sum(getattr(ds._data, k)[lo:hi].numel() * 4 for k in ds.keys) + plan.buf.numel()    # spectra + plan tables, per rank
    d2h = Bsz * world * record_bytes(N_PAD) + n_atoms.nbytes
    if args.cpu_baseline:
        cpu_rate, cpu_s, cpu_desc = cpu_reference_rate(16, 2, model=w['model'], n_pad=N_PAD, all_max=w['all_max'])
        c0 = cpu_configs0(repeats=args.cpu_repeats) if (w['model'] == 'DMT' and N_PAD == 29 and args.cpu_repeats > 0) else None
    else:
        cpu_rate, cpu_desc, c0 = None, 'skipped (--no-cpu-baseline)', None
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'strong' if w['strong'] else 'weak', 'vs_baseline': None,
        'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args, w), 'model': w['model'] + '_B200',
                   'batch_per_gpu': Bsz, 'molecules_per_step': n_items, 'diffusion_steps': S, 'n_pad': N_PAD,
                   'n_atoms': 'all %d' % N_PAD if w['all_max'] else 'QM9S histogram, mean %.2f' % n_atoms.mean(),
                   'noise': 'device Philox', 'l2': 'inputs_exceed_l2 (per-step working set >> 126 MB)',
                   'parallelism': 'dp%d independent shards + 1 all-gather/round' % world},
        'denoiser_steps_per_s': args.steps * S / (ms / 1000.0),
        'molecule_steps_per_s': value * S,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'ms_per_step': ms_e2e / args.steps, 'ms_per_step_device': ms_e2e_dev / args.steps,
                'api': 'get_cond_sampling_eval_fn(config, ...)(model) on an in-memory data set with pinned host spectra'},
        'gpu_launches': int(launches),
        'clocks': clk,
        'roofline': dict(kern['dominant'], step={
            'bound': 'tensor', 'achieved': achieved, 'peak': peaks['sustained'], 'unit': 'TFLOP/s', 'frac': achieved / peaks['sustained'],
            'what': 'whole denoiser step: algorithmic FLOPs per molecule-step (SURVEY.md 8(d); bench.alg_flops) / CUDA-event time '
                    'of the timed rounds; peak = sustained bf16, ' + peaks['which']}, kernels=kern['kernels'],
            in_stream_step_us=kern['step_us']),
        'cpu_baseline': {'value': cpu_rate, 'unit': UNIT, 'cores': host_threads(), 'kind': 'port', 'sample': cpu_desc, 'configs0': c0},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _demangle(name):
    import re
    m = re.search(r'\d+(k_[a-z0-9_]+|gemm_tc_kernel|gemm_simt_kernel|coord_head_kernel|edge_ffn_kernel)', name)
    base = m.group(1) if m else name[:40]
    t = re.search(r'gemm_tc_kernelILi(\d+)ELi(\d+)ELb(\d)E', name)
    if t:
        base += '<%s,%s,%s>' % (t.group(1), ('STORE', 'LNMOD', 'RESGATE', 'COORD', 'EHEAD')[int(t.group(2))], t.group(3))
    return base


def csrc_sha():
    """sha256 over the CUDA sources + the header: identifies the build a profile was taken from (the GPU box has no .git)."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, 'diffspectra_b200', 'csrc')
    for f in sorted(os.listdir(d)) + ['../../include/diffspectra_b200.h']:
        h.update(f.encode())
        h.update(open(os.path.join(d, f), 'rb').read())
    return h.hexdigest()[:16]


def ncu_dram_table():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
    (profiles/ncu_dram.json, written by scratch/ncu_dram.py from the .ncu-rep of `bench.py --workload configs1`).  Only
    returned when the capture was taken from THIS build (same csrc_sha): a stale table is never reported as traffic."""
    p = os.path.join(ROOT, 'profiles', 'ncu_dram.json')
    if not os.path.isfile(p):
        return {}, 'no capture (profiles/ncu_dram.json missing)'
    d = json.load(open(p))
    if d.get('csrc_sha') != csrc_sha():
        return {}, 'capture %s is from another build (csrc_sha %s != %s)' % (d.get('source'), d.get('csrc_sha'), csrc_sha())
    return d.get('kernels', {}), '%s (csrc_sha %s)' % (d.get('source'), d.get('csrc_sha'))


def step_kernel_profile(eng, plan, coef, spectra_dev, n_atoms, args, w):
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
    wo = w['model'] != 'DMT'
    Me = Md if wo else Mp                       # rows of the edge tensors
    bytes_of = {      # ALGORITHMIC bytes per launch of the graph-side kernels (what must cross HBM once)
        'k_attention_grp': Mp * 1024 + Mp + Mn * (1536 + 256 * 6),          # e0|e1 once per pair, flags, q|k|v, hn fp32+bf16
        'k_wo_attention': Md * 1024 + Mn * (1536 + 512),
        'k_coord_ln': Mp * 512 + Mn * 1024 + Md * 513,                       # gp, ab in; Z + flags out
        'coord_head_kernel': Mp * (256 + 1) + Mn * 1024 + Md * 4,            # X in, pair flags, per-atom table ab (L2-resident), w out
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
            if kname.startswith('coord_head_kernel'):      # pair part of input_lin once per pair + coord_mlp.0 per directed edge
                flops = 2.0 * Mp * 128 * 256 + 2.0 * Md * 256 * 256
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
    dram, dram_src = ncu_dram_table()
    for r in rows:
        if args.workload == 'configs1' and r['kernel'] in dram:
            r['ncu_dram_bytes'] = dram[r['kernel']]
    dominant = {
        'kernel': top['kernel'] + (' ' + top['shape'] if 'shape' in top else ''),
        'bound': 'tensor' if tensor_bound else 'hbm',
        'achieved': top.get('tflops') if tensor_bound else top.get('gbs'),
        'peak': peaks['burst'] if tensor_bound else peaks['hbm'],
        'unit': 'TFLOP/s' if tensor_bound else 'GB/s',
        'frac': top.get('frac_tensor') if tensor_bound else top.get('frac_hbm'),
        'traffic': (dram[top['kernel']] if (args.workload == 'configs1' and w['tag'] == WORKLOADS['configs1']['tag'] and top['kernel'] in dram)
                    else None),
        'traffic_source': dram_src,
        'us_per_launch': top['us_per_launch'], 'share_of_step': top['share_of_step'],
        'what': 'dominant kernel of a denoiser step by in-stream CUDA-event time; achieved = algorithmic bytes (or FLOPs) per '
                'launch / average launch time; peak = ' + peaks['which'] + ' (burst figures); traffic = ncu dram bytes per launch '
                '(profiles/ncu_dram.json, only when taken from this build), bytes',
    }
    return {'dominant': dominant, 'kernels': rows[:14], 'step_us': total / steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='configs1', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=None, help='override the per-GPU batch of the workload')
    ap.add_argument('--diffusion-steps', type=int, default=1000)
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--model', default=None, choices=['DMT', 'DMT_WO_EQ'], help='override the model of the workload')
    ap.add_argument('--n-pad', type=int, default=None, help='override the padded atoms per molecule (> 29 implies --all-max)')
    ap.add_argument('--all-max', '--all29', dest='all_max', action='store_true', help='worst case: every molecule has n_pad atoms')
    ap.add_argument('--no-cpu-baseline', dest='cpu_baseline', action='store_false')
    ap.add_argument('--cpu-repeats', type=int, default=3, help='full runs of BASELINE configs[0] on the host cores (0 = skip)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
