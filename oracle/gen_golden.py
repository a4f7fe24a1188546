"""Generate tests/golden/*.pt from the UNMODIFIED reference (build container only).

    python -m oracle.gen_golden            # writes tests/golden/

Every fixture is produced by reference code (models/dmt.py, models/specformer.py, sampling.py,
diffusion/noise_schedule.py, utils.py) imported through oracle/ref_harness.py, on name-keyed
deterministic weights (oracle/weights.py) so the 160 MB state_dict is NOT stored — only inputs
that cannot be regenerated from a seed and the reference OUTPUTS are.  TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import sys

import torch

from . import weights as W
from .ref_harness import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def _model(R, version, salt=0, coord_scale=None, cls='DMT'):
    R.config.data.spectra_version = version
    m = getattr(R, cls)(R.config).eval()
    sd = m.state_dict()
    W.keyed_fill_(sd, salt=salt, coord_scale=coord_scale)
    m.load_state_dict(sd)
    return m


def _sym_edge(B, N, C, em, g, scale=1.0):
    z = torch.randn(B, C, N, N, generator=g) * scale
    z = torch.tril(z, -1)
    z = z + z.transpose(-1, -2)
    return z.permute(0, 2, 3, 1) * em.reshape(B, N, N, 1)


def gen_manifest(R):
    out = {}
    for version in ('allspectra', 'ir'):
        m = _model(R, version)
        out[version] = {
            'params': [[n, list(p.shape)] for n, p in m.named_parameters()],
            'buffers': [[n, list(b.shape)] for n, b in m.named_buffers()],
        }
    with open(os.path.join(OUT, 'param_manifest.json'), 'w') as f:
        json.dump(out, f)


def gen_schedule(R):
    ns = R.NoiseScheduleVP('cosine', continuous_beta_0=0.1, continuous_beta_1=20.)
    res = {}
    for steps in (50, 1000):
        t_array = torch.linspace(ns.T, 1e-3, steps)
        s_array = torch.cat([t_array[1:], torch.zeros(1)])
        rows = []
        for i in range(steps):                          # sampling.py:571-584,605-606
            t, s = t_array[i], s_array[i]
            alpha_t, sigma_t = ns.marginal_prob(t)
            alpha_s, sigma_s = ns.marginal_prob(s)
            a_ts = alpha_t / alpha_s
            s2 = sigma_t ** 2 - a_ts ** 2 * sigma_s ** 2
            sigma = torch.sqrt(s2) * sigma_s / sigma_t
            rows.append(torch.stack([a_ts * sigma_s ** 2 / sigma_t ** 2, alpha_s * s2 / sigma_t ** 2, sigma,
                                     torch.log(alpha_t ** 2 / sigma_t ** 2)]))
        res['table_%d' % steps] = torch.stack(rows)
    torch.save(res, os.path.join(OUT, 'schedule.pt'))


def gen_denoiser(R):
    """Teacher-forced single denoiser calls: step 0 (no cond), a self-conditioned step and a step with
    adversarial cond inputs (mixed-sign cond edges, spread cond positions) so both adjacency heads vary."""
    for version, salt, cs in (('allspectra', 0, None), ('ir', 1, 0.05)):
        m = _model(R, version, salt=salt, coord_scale=cs)
        n = torch.tensor([29, 9, 17, 23, 3, 12]) if version == 'allspectra' else torch.tensor([20, 29, 5, 14])
        B, N = len(n), 29
        nm, em = W.make_masks(n, N)
        ctx = W.synthetic_spectra(B, version, seed=77 + salt)
        g = torch.Generator().manual_seed(100 + salt)
        from .dense_oracle import node_noise_from_raw
        x = node_noise_from_raw(torch.randn(B, N, 3, generator=g), torch.randn(B, N, 6, generator=g), nm)
        ex = _sym_edge(B, N, 2, em, g)
        cases = {}
        with torch.no_grad():
            nl0 = torch.full((B,), -9.5)
            p0, e0 = m(nl0, x, nm, em, edge_x=ex, noise_level=nl0, cond_x=None, cond_edge_x=None, context=ctx)
            cases['step0'] = dict(x=x, edge_x=ex, noise_level=nl0, cond_x=None, cond_edge_x=None, pred=p0, edge_pred=e0)
            x1, ex1 = 0.9 * x + 0.1 * p0, 0.9 * ex + 0.1 * e0
            nl1 = torch.linspace(-3., 4., B)             # per-molecule noise levels (training-style call)
            p1, e1 = m(nl1, x1, nm, em, edge_x=ex1, noise_level=nl1, cond_x=p0, cond_edge_x=e0, context=ctx)
            cases['selfcond'] = dict(x=x1, edge_x=ex1, noise_level=nl1, cond_x=p0, cond_edge_x=e0, pred=p1, edge_pred=e1)
            cx = node_noise_from_raw(torch.randn(B, N, 3, generator=g) * 1.2, torch.randn(B, N, 6, generator=g) * 0.3, nm)
            cex = _sym_edge(B, N, 2, em, g, 0.5)
            nl2 = torch.full((B,), 6.0)
            p2, e2 = m(nl2, x1, nm, em, edge_x=ex1, noise_level=nl2, cond_x=cx, cond_edge_x=cex, context=ctx)
            cases['advcond'] = dict(x=x1, edge_x=ex1, noise_level=nl2, cond_x=cx, cond_edge_x=cex, pred=p2, edge_pred=e2)
            cemb = m.cond_lin(m.cond_encoder(ctx))
            spec = m.cond_encoder(ctx)
        torch.save(dict(version=version, salt=salt, coord_scale=cs, n_atoms=n, N=N, ctx_seed=77 + salt,
                        cases=cases, ctx_emb=cemb, spec_emb=spec),
                   os.path.join(OUT, 'denoiser_%s.pt' % version))


def gen_denoiser_wo_eq(R):
    """DMT_WO_EQ (models/dmt_wo_eq.py): teacher-forced calls at step 0 and a self-conditioned step, plus a 10-step
    free-running trajectory with the reference sampler, allspectra."""
    from . import dense_oracle as O
    version, salt = 'allspectra', 3
    m = _model(R, version, salt=salt, cls='DMT_WO_EQ')
    n = torch.tensor([29, 9, 17, 2, 23])
    B, N = len(n), 29
    nm, em = W.make_masks(n, N)
    ctx = W.synthetic_spectra(B, version, seed=91)
    g = torch.Generator().manual_seed(300)
    x = O.node_noise_from_raw(torch.randn(B, N, 3, generator=g), torch.randn(B, N, 6, generator=g), nm)
    ex = _sym_edge(B, N, 2, em, g)
    cases = {}
    with torch.no_grad():
        nl0 = torch.full((B,), -9.5)
        p0, e0 = m(nl0, x, nm, em, edge_x=ex, noise_level=nl0, cond_x=None, cond_edge_x=None, context=ctx)
        cases['step0'] = dict(x=x, edge_x=ex, noise_level=nl0, cond_x=None, cond_edge_x=None, pred=p0, edge_pred=e0)
        x1, ex1 = 0.9 * x + 0.1 * p0, 0.9 * ex + 0.1 * e0
        nl1 = torch.linspace(-3., 4., B)
        p1, e1 = m(nl1, x1, nm, em, edge_x=ex1, noise_level=nl1, cond_x=p0, cond_edge_x=e0, context=ctx)
        cases['selfcond'] = dict(x=x1, edge_x=ex1, noise_level=nl1, cond_x=p0, cond_edge_x=e0, pred=p1, edge_pred=e1)
        ns = R.NoiseScheduleVP('cosine', continuous_beta_0=0.1, continuous_beta_1=20.)
        steps = 10
        sampler = R.AncestralSampler(ns, torch.linspace(ns.T, 1e-3, steps), True, True, True,
                                     R.get_self_cond_fn(R.config), sampling_temperature=1.0)
        torch.manual_seed(43)
        z = R.mutils.sample_combined_position_feature_noise(B, N, 6, nm)
        ez = R.mutils.sample_symmetric_edge_feature_noise(B, N, 2, em)
        x_mean, ex_mean = sampler.sampling(m, z, nm, em, ez, ctx)
    torch.save(dict(version=version, salt=salt, coord_scale=None, n_atoms=n, N=N, ctx_seed=91, cases=cases,
                    traj=dict(steps=steps, seed=43, x_mean=x_mean, edge_x_mean=ex_mean)),
               os.path.join(OUT, 'denoiser_wo_eq_allspectra.pt'))


def gen_sampler(R, full=False):
    """Free-running ancestral sampling with the reference AncestralSampler + post_process."""
    from . import dense_oracle as O
    ns = R.NoiseScheduleVP('cosine', continuous_beta_0=0.1, continuous_beta_1=20.)
    cases = (('ir', 50, torch.tensor([29, 11, 18, 22]), 1), ('allspectra', 20, torch.tensor([16, 29, 7]), 0))
    if full:     # the shipped sampling length (configs/diffspectra_qm9s.py:147: 1000 steps), a few minutes of CPU
        cases = (('allspectra', 1000, torch.tensor([29, 9, 17, 23]), 2),)
    for version, steps, n, salt in cases:
        m = _model(R, version, salt=salt)
        B, N = len(n), 29
        nm, em = W.make_masks(n, N)
        ctx = W.synthetic_spectra(B, version, seed=55)
        sampler = R.AncestralSampler(ns, torch.linspace(ns.T, 1e-3, steps), True, True, True,
                                     R.get_self_cond_fn(R.config), sampling_temperature=1.0)
        torch.manual_seed(42)
        z = R.mutils.sample_combined_position_feature_noise(B, N, 6, nm)
        ez = R.mutils.sample_symmetric_edge_feature_noise(B, N, 2, em)
        with torch.no_grad():
            x_mean, ex_mean = sampler.sampling(m, z, nm, em, ez, ctx)
            inv = R.get_data_inverse_scaler(R.config)
            pos, one_hot, fc, bond = R.post_process(x_mean, 5, True, nm, inv, ex_mean, em, True)
        mols = R.mol_process(one_hot, pos, fc, [int(v) for v in n], bond)
        # one fixed-noise sampler-step KAT (sampling.py:605-624) taken from an independent draw
        torch.save(dict(version=version, steps=steps, n_atoms=n, N=N, salt=salt, seed=42, ctx_seed=55,
                        x_mean=x_mean, edge_x_mean=ex_mean, pos=pos, one_hot=one_hot, fc=fc, bond=bond,
                        mols=[(a, b.long(), c, d.long()) for a, b, c, d in mols]),
                   os.path.join(OUT, 'sampler_%s_%d.pt' % (version, steps)))


def gen_sampler_stat(R, B=128, steps=50, salt=4, seed=42, ctx_seed=56, n_seed=4321, tag=None):
    """Statistics-sized free-running trajectory (north_star: ">= 99 % of molecules with identical atom types and bond
    orders, RMSD <= 1e-3 A"): B molecules with QM9S-histogram atom counts, allspectra, the reference AncestralSampler +
    post_process.  Stored compactly: the final means (fp32) and the discrete molecules as small integers."""
    ns = R.NoiseScheduleVP('cosine', continuous_beta_0=0.1, continuous_beta_1=20.)
    version = 'allspectra'
    m = _model(R, version, salt=salt)
    n = W.sample_n_atoms(B, seed=n_seed)
    N = 29
    nm, em = W.make_masks(n, N)
    ctx = W.synthetic_spectra(B, version, seed=ctx_seed)
    sampler = R.AncestralSampler(ns, torch.linspace(ns.T, 1e-3, steps), True, True, True,
                                 R.get_self_cond_fn(R.config), sampling_temperature=1.0)
    torch.manual_seed(seed)
    z = R.mutils.sample_combined_position_feature_noise(B, N, 6, nm)
    ez = R.mutils.sample_symmetric_edge_feature_noise(B, N, 2, em)
    with torch.no_grad():
        x_mean, ex_mean = sampler.sampling(m, z, nm, em, ez, ctx)
        inv = R.get_data_inverse_scaler(R.config)
        pos, one_hot, fc, bond = R.post_process(x_mean, 5, True, nm, inv, ex_mean, em, True)
    iu = torch.triu_indices(N, N, 1)
    torch.save(dict(version=version, steps=steps, n_atoms=n, N=N, salt=salt, seed=seed, ctx_seed=ctx_seed,
                    x_mean=x_mean, edge_x_mean_triu=ex_mean[:, iu[0], iu[1]].contiguous(), pos=pos,
                    atom=one_hot.argmax(-1).to(torch.uint8), fc=fc.squeeze(-1).to(torch.int8), bond=bond.to(torch.uint8)),
               os.path.join(OUT, tag or 'sampler_stat_%s_%d_b%d.pt' % (version, steps, B)))


def gen_noise_kat(R):
    n = torch.tensor([29, 4, 13])
    B, N = 3, 29
    nm, em = W.make_masks(n, N)
    torch.manual_seed(7)
    z = R.mutils.sample_combined_position_feature_noise(B, N, 6, nm)
    ez = R.mutils.sample_symmetric_edge_feature_noise(B, N, 2, em)
    torch.save(dict(n_atoms=n, N=N, seed=7, z=z, edge_z=ez), os.path.join(OUT, 'noise_kat.pt'))


def main():
    os.makedirs(OUT, exist_ok=True)
    R = load_reference()
    if '--only-full' in sys.argv:
        gen_sampler(R, full=True)
        return
    if '--only-stat' in sys.argv:        # B=128 x 50 steps (a few minutes of CPU)
        gen_sampler_stat(R)
        return
    if '--only-stat-full' in sys.argv:   # B=32 x 1000 steps (about an hour of CPU)
        gen_sampler_stat(R, B=32, steps=1000, salt=5, seed=43, ctx_seed=57, n_seed=4322)
        return
    gen_manifest(R)
    gen_schedule(R)
    gen_noise_kat(R)
    gen_denoiser(R)
    gen_denoiser_wo_eq(R)
    gen_sampler(R)
    if '--full' in sys.argv:
        gen_sampler(R, full=True)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()
