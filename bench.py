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
N_PAD = 29


def alg_flops(n):
    """ALGORITHMIC FLOPs of one molecule x one denoiser step with n atoms."""
    return 2 * (21007360 + 5197568 * n + 1290560 * n * (n - 1))


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
def cpu_reference_rate(sample_b=16, sample_steps=4, repeats=1):
    """Oracle port of the reference PyTorch path (dense restatement, oracle/dense_oracle.py) on the host cores,
    allspectra, on `sample_b` molecules x `sample_steps` denoiser steps (+ one SpecFormer pass per step, as the
    reference recomputes it every call), extrapolated to 1000 steps.  Returns (molecules/s, seconds, description)."""
    import torch
    from diffspectra_b200.config import get_config
    from diffspectra_b200.model import DMT_B200
    from oracle import dense_oracle as O
    from oracle import weights as W
    torch.manual_seed(42)
    sd = DMT_B200(get_config(VERSION, device='cpu')).state_dict()
    n = W.sample_n_atoms(sample_b, seed=1234)
    nm, em = W.make_masks(n, N_PAD)
    ctx = W.synthetic_spectra(sample_b, VERSION, seed=1235)
    table = O.schedule_table(1000)[:: max(1, 1000 // sample_steps)][:sample_steps]
    g = torch.Generator().manual_seed(42)
    z = O.node_noise_from_raw(torch.randn(sample_b, N_PAD, 3, generator=g), torch.randn(sample_b, N_PAD, 6, generator=g), nm)
    ez = O.edge_noise_from_raw(torch.randn(sample_b, 2, N_PAD, N_PAD, generator=g), em)
    raw = [O.draw_step_noise(sample_b, N_PAD, nm, em, generator=g) for _ in range(sample_steps)]

    def denoise(x, ex, nl, cx, cex):          # the reference re-runs SpecFormer inside every call (dmt.py:348-350)
        return O.dmt_forward(sd, x, nm, em, ex, nl, cx, cex, O.context_embedding(sd, ctx, VERSION))

    best = None
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            O.ancestral_sample(denoise, table, z, ez, nm, em, raw)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    per_step = best / sample_steps
    rate = sample_b / (per_step * 1000.0)
    desc = ('oracle port of the reference PyTorch path, allspectra, B=%d molecules (QM9S histogram, N_pad=29) x %d of 1000 '
            'denoiser steps (SpecFormer recomputed each step like the reference), %.2f s/step, extrapolated to 1000 steps'
            % (sample_b, sample_steps, per_step))
    return rate, best, desc


def run_reference(args):
    import torch
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = torch.get_num_threads()
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_rate(8, 1)
    rates, secs = [], 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r, s, desc = cpu_reference_rate(16, 2)
        rates.append(r)
        secs += s
        if time.perf_counter() - t0 > 150:
            break
    value = len(rates) * 16 / sum(16 / r for r in rates)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': len(rates),
        'warmup': min(args.warmup, 1), 'ms_per_step': 1000.0 * 1024 / value, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'DMT allspectra sampling, 1000 steps, batch 1024 (BASELINE.json configs[1]); bounded sample',
                   'batch': 1024, 'diffusion_steps': 1000},
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
    from diffspectra_b200.model import DMT_B200
    from diffspectra_b200.noise_schedule import NoiseScheduleVP, ancestral_coefficients
    from oracle import weights as W          # synthetic inputs only (atom-count histogram, spectra)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if rank == 0:
        B.build()
    if world > 1:
        dist.barrier()

    Bsz, S = args.batch, args.diffusion_steps
    torch.manual_seed(42)                                   # random-init weights of the reference architecture
    model = DMT_B200(get_config(VERSION, device=str(dev), precision=args.precision)).eval().to(dev)
    eng = model.engine(dev)
    n_atoms = W.sample_n_atoms(Bsz, seed=1234 + rank).numpy().astype(np.int32)
    if args.all29:
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
        kern = kernel_rooflines(eng, dev) if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    mols = Bsz * world * args.steps
    value = mols / (ms / 1000.0)
    e2e_value = mols / (ms_e2e / 1000.0)
    flops_round = float(sum(alg_flops(int(n)) for n in n_atoms)) * S            # rank 0's shard, denoiser only
    achieved = flops_round * args.steps / (ms / 1000.0) / 1e12                  # TFLOP/s per GPU
    h2d = sum(t.numel() * 4 for t in spectra_host) + n_atoms.nbytes
    d2h = int(rec_host.numel() * rec_host.element_size()) if rec_host is not None else 0
    cpu_rate, cpu_s, cpu_desc = cpu_reference_rate(16, 2) if args.cpu_baseline else (None, 0, 'skipped (--no-cpu-baseline)')
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
        'config': {'workload': 'DMT allspectra sampling, %d steps, batch %d per GPU (BASELINE.json configs[1])' % (S, Bsz),
                   'batch_per_gpu': Bsz, 'diffusion_steps': S, 'n_atoms': 'all 29' if args.all29 else 'QM9S histogram, mean %.2f' % n_atoms.mean(),
                   'noise': 'device Philox', 'l2': 'inputs_exceed_l2 (per-step working set >> 126 MB)',
                   'parallelism': 'dp%d independent shards + 1 all-gather/round' % world},
        'denoiser_steps_per_s': args.steps * S / (ms / 1000.0),
        'molecule_steps_per_s': value * S,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': d2h,
                'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': int(launches),
        'clocks': clk,
        'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peaks['sustained'], 'unit': 'TFLOP/s',
                     'frac': achieved / peaks['sustained'], 'traffic': None,
                     'what': 'whole denoiser step: algorithmic FLOPs 2*(21007360+5197568 n+1290560 n(n-1)) per molecule-step '
                             '(SURVEY.md 8(d)) / CUDA-event time of the timed rounds; peak = sustained bf16, ' + peaks['which'],
                     'kernels': kern},
        'cpu_baseline': {'value': cpu_rate, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port', 'sample': cpu_desc},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def kernel_rooflines(eng, dev):
    """The dominant kernel timed ALONE with CUDA events: the tcgen05 GEMM on the two largest contraction shapes of a
    step (coord_mlp.0 over directed edges, lin_edge0|lin_edge1 over pairs), against the burst bf16 peak."""
    import ctypes
    import torch
    from diffspectra_b200 import _lib as L
    peaks = load_peaks()
    res = []
    for name, M, N, K, act in (('coord_mlp.0 [2Mp,256]x[256,256]+SiLU', 323072, 256, 256, L.ACT_SILU),
                               ('lin_edge0|1 [Mp,64]x[512,64]+tanh', 161536, 512, 64, L.ACT_TANH),
                               ('adaLN table [B,1024]x[19584,1024]', 1024, 19584, 1024, L.ACT_NONE)):
        A = torch.randn(M, K, device=dev).bfloat16()
        Wt = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
        bias = torch.zeros(N, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)

        def go():
            L.check(L.lib().ds_gemm(eng.h, 1, L.ptr(A), K, L.ptr(Wt), K, L.ptr(bias), ctypes.c_void_p(0), 0, L.ptr(out), N,
                                    M, N, K, L.DT_BF16, L.DT_BF16, act, L.stream_ptr()), 'ds_gemm')
        for _ in range(3):
            go()
        flush = torch.empty(64 << 20, device=dev, dtype=torch.float32)
        ts = []
        for _ in range(5):
            flush.fill_(1.0)                              # L2 flush between timed launches (256 MB > 126 MB L2)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); go(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        tf = 2.0 * M * N * K / (ms / 1e3) / 1e12
        gbs = (M * K * 2 + N * K * 2 + M * N * 2) / (ms / 1e3) / 1e9
        res.append({'kernel': 'gemm_tc_kernel', 'shape': name, 'ms': ms, 'tflops': tf, 'frac_of_burst_bf16': tf / peaks['burst'],
                    'gbs': gbs, 'frac_of_hbm': gbs / peaks['hbm']})
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=1024)
    ap.add_argument('--diffusion-steps', type=int, default=1000)
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--all29', action='store_true', help='worst case: every molecule has 29 atoms')
    ap.add_argument('--no-cpu-baseline', dest='cpu_baseline', action='store_false')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
