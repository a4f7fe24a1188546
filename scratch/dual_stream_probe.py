"""Probe: 1024 molecules as ONE batch vs as K concurrent sub-batches (own context / plan / CUDA graph / stream each).
Same kernels; the question is whether kernels of different sub-batches fill each other's ramp + tail."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from diffspectra_b200.config import get_config
from diffspectra_b200.engine import Engine
from diffspectra_b200.model import DMT_B200
from diffspectra_b200.noise_schedule import NoiseScheduleVP, ancestral_coefficients
from oracle import weights as W

dev = torch.device('cuda', 0)
Bsz, S, N = 1024, int(os.environ.get('STEPS', '100')), 29
torch.manual_seed(42)
model = DMT_B200(get_config('allspectra', device=str(dev), precision='bf16')).eval().to(dev)
sd = model.state_dict()
n_atoms = W.sample_n_atoms(Bsz, seed=1234, max_n=29).numpy().astype(np.int32)
spectra = [t.to(dev) for t in W.synthetic_spectra(Bsz, 'allspectra', seed=1235)]
ns = NoiseScheduleVP('cosine', continuous_beta_0=0.1, continuous_beta_1=20.)
coef = ancestral_coefficients(ns, torch.linspace(ns.T, 1e-3, S, device=dev))

def setup(K):
    parts = []
    per = Bsz // K
    for k in range(K):
        eng = Engine(dev, mode='bf16', spectra_version='allspectra', model_kind='DMT')
        eng.pack_weights(sd)
        sl = slice(k * per, (k + 1) * per)
        plan = eng.plan(n_atoms[sl], N)
        sp = [t[sl].contiguous() for t in spectra]
        out = (torch.empty(per, N, 9, device=dev), torch.empty(per, N, N, 2, device=dev))
        parts.append(dict(eng=eng, plan=plan, sp=sp, out=out, stream=torch.cuda.Stream(), gid=k * per))
    return parts

def run(parts):
    cur = torch.cuda.current_stream()
    for p in parts:
        p['stream'].wait_stream(cur)
    for p in parts:
        with torch.cuda.stream(p['stream']):
            ctx = p['eng'].context_embedding(p['sp'])
            p['eng'].sample_loop(p['plan'], ctx, coef, None, None, None, seed=42, gid_base=p['gid'], temperature=1.0,
                                 use_graph=True, out=p['out'])
    for p in parts:
        cur.wait_stream(p['stream'])

with torch.no_grad():
    ref = None
    for K in (1, 2, 4):
        parts = setup(K)
        for _ in range(2):
            run(parts)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            run(parts)
        e1.record()
        torch.cuda.synchronize()
        x = torch.cat([p['out'][0] for p in parts])
        if ref is None:
            ref = x.clone()
        print('K=%d  %.1f ms per round of %d steps   max |x - x(K=1)| = %.3e' % (K, e0.elapsed_time(e1) / 3, S, (x - ref).abs().max().item()), flush=True)
        del parts
        torch.cuda.empty_cache()
