"""Parity of the sampler-side CUDA path (ds_sampler_step, ds_sample_loop, ds_post_process) through the
reference-facing AncestralSampler / post_process, against reference goldens and the oracle."""
import numpy as np
import pytest
import torch

from oracle import dense_oracle as O
from oracle import philox_ref as P
from oracle import weights as W
from diffspectra_b200.noise_schedule import NoiseScheduleVP
from tests.helpers import build_model, load_golden, max_abs, rel_l2

pytestmark = pytest.mark.gpu


def _sampler(steps, device='cpu', **kw):
    """time_steps on the CPU: the coefficient table is then computed with the same libm as the CPU-generated reference
    goldens (on the GPU the last row differs: alpha_s rounds to exactly 1 there)."""
    from diffspectra_b200.noise_schedule import NoiseScheduleVP
    from diffspectra_b200.sampling import AncestralSampler
    ns = NoiseScheduleVP('cosine', continuous_beta_0=0.1, continuous_beta_1=20.)
    return AncestralSampler(ns, torch.linspace(ns.T, 1e-3, steps, device=device), True, True, True, None, 1.0, **kw)


def test_coefficient_table_matches_reference_golden():
    g = load_golden('schedule.pt')
    for steps in (50, 1000):
        ref = g['table_%d' % steps]
        tab = _sampler(steps).coefficients()
        assert torch.allclose(tab, ref, rtol=1e-6, atol=1e-9)
        # computed on the GPU (what the reference does in production): ulp-level libm differences are amplified by
        # 1 - exp(2 log alpha) in the last row (alpha_s == 1 exactly on the GPU), elsewhere they stay small
        tab_gpu = _sampler(steps, device='cuda').coefficients().cpu()
        err = ((tab_gpu[:-1] - ref[:-1]).abs() / ref[:-1].abs().clamp(min=1e-6)).max().item()
        print(steps, 'GPU-vs-CPU table max rel err (all but last row)', err, 'last row', tab_gpu[-1].tolist(), ref[-1].tolist())
        assert err < 5e-3


def test_sampler_step_matches_oracle():
    """One fused ancestral update with the reference's randn draws (sampling.py:605-624)."""
    n = torch.tensor([29, 4, 13, 1, 2])
    B, N = len(n), 29
    nm, em = W.make_masks(n, N)
    model = build_model('ir', precision='fp32')
    eng = model.engine('cuda')
    plan = eng.plan(n.numpy(), N)
    g = torch.Generator().manual_seed(3)
    x = O.node_noise_from_raw(torch.randn(B, N, 3, generator=g), torch.randn(B, N, 6, generator=g), nm)
    ex = O.edge_noise_from_raw(torch.randn(B, 2, N, N, generator=g), em)
    pred = O.node_noise_from_raw(torch.randn(B, N, 3, generator=g), torch.randn(B, N, 6, generator=g), nm)
    epred = O.edge_noise_from_raw(torch.randn(B, 2, N, N, generator=g), em)
    rp, rh, re = O.draw_step_noise(B, N, nm, em, generator=g)
    row = torch.tensor([0.9954, 0.00437, 0.0556, -0.0078])
    T = 0.8
    x_mean = row[0] * x + row[1] * pred
    x_new = x_mean + row[2] * O.node_noise_from_raw(rp, rh, nm) * T
    e_mean = row[0] * ex + row[1] * epred
    e_new = e_mean + row[2] * O.edge_noise_from_raw(re, em) * T
    ox, oe, oxm, oem = eng.sampler_step(plan, x, ex, pred, epred, row, (rp, rh, re), temperature=T)
    for a, b in ((ox, x_new), (oe, e_new), (oxm, x_mean), (oem, e_mean)):
        assert max_abs(a, b) < 2e-6
    # padding stays exactly zero and the coordinate noise is centre-of-mass free
    assert (ox.cpu() * (1 - nm)).abs().max() == 0
    assert ((ox.cpu() - oxm.cpu())[..., :3].sum(1).abs().max()) < 1e-5


def test_noise_kat_vs_reference_golden():
    """The masked / CoM-free / mirrored noise construction (models/utils.py:67-106) from raw randn draws."""
    g = load_golden('noise_kat.pt')
    n, N = g['n_atoms'], g['N']
    B = len(n)
    nm, em = W.make_masks(n, N)
    torch.manual_seed(g['seed'])
    rp, rh, re = torch.randn(B, N, 3), torch.randn(B, N, 6), torch.randn(B, 2, N, N)
    model = build_model('ir', precision='fp32')
    eng = model.engine('cuda')
    plan = eng.plan(n.numpy(), N)
    zero = torch.zeros(B, N, 9)
    ezero = torch.zeros(B, N, N, 2)
    row = torch.tensor([0., 0., 1., 0.])            # x <- 1 * noise
    ox, oe, _, _ = eng.sampler_step(plan, zero, ezero, zero, ezero, row, (rp, rh, re))
    assert max_abs(ox, g['z']) < 1e-6
    assert max_abs(oe, g['edge_z']) == 0


@pytest.mark.parametrize('name,use_graph', [('sampler_ir_50.pt', True), ('sampler_ir_50.pt', False),
                                            ('sampler_allspectra_20.pt', True),
                                            ('sampler_allspectra_1000.pt', True)])     # the shipped length: 1000 steps
def test_free_running_sampling_matches_reference_golden(name, use_graph):
    """Whole loop, fp32 mode, the reference's own RNG stream (torch.manual_seed(42) + its draw order):
    final argmax atom types / bond orders identical, coordinate RMSD <= 1e-3 A (north_star)."""
    from diffspectra_b200.sampling import make_masks, post_process
    g = load_golden(name)
    n, N, version = g['n_atoms'], g['N'], g['version']
    B = len(n)
    model = build_model(version, g['salt'], None, 'fp32')
    nm, em = make_masks(n, 'cuda', N)
    ctx = W.synthetic_spectra(B, version, seed=g['ctx_seed'])
    sampler = _sampler(g['steps'], noise='torch', use_graph=use_graph)
    # the reference draws on the CPU generator; replay the same stream there and hand the draws over
    torch.manual_seed(g['seed'])
    nm_c, em_c = W.make_masks(n, N)
    z = O.node_noise_from_raw(torch.randn(B, N, 3), torch.randn(B, N, 6), nm_c)
    ez = O.edge_noise_from_raw(torch.randn(B, 2, N, N), em_c)
    raw = [O.draw_step_noise(B, N, nm_c, em_c) for _ in range(g['steps'])]
    eng = model.engine('cuda')
    plan = model.plan_for(nm)
    ctx_emb = model.context_embedding(ctx)
    rp = torch.stack([r[0] for r in raw]); rh = torch.stack([r[1] for r in raw]); re = torch.stack([r[2] for r in raw])
    x_mean, e_mean = eng.sample_loop(plan, ctx_emb, sampler.coefficients(), z, ez, (rp, rh, re), use_graph=use_graph)
    print(name, 'x_mean rel %.2e edge rel %.2e' % (rel_l2(x_mean, g['x_mean']), rel_l2(e_mean, g['edge_x_mean'])))
    pos, one_hot, fc, bond = post_process(x_mean, 5, True, nm, None, e_mean, em, True, model=model)
    rmsd = ((pos.cpu() - g['pos']) ** 2).sum(-1).sum(-1).div(n.float()).sqrt()
    print('rmsd per molecule', rmsd.tolist())
    assert rmsd.max().item() <= 1e-3
    assert torch.equal(one_hot.cpu().argmax(-1), g['one_hot'].argmax(-1))
    assert torch.equal(fc.cpu().long(), g['fc'].long())
    assert torch.equal(bond.cpu(), g['bond'])


def test_segmented_loop_equals_single_call():
    """noise='torch' draws noise in segments; any segmentation must give the same trajectory."""
    from diffspectra_b200.sampling import AncestralSampler, make_masks
    n = torch.tensor([12, 29, 3])
    B, N = 3, 29
    model = build_model('ir', precision='fp32')
    nm, em = make_masks(n, 'cuda', N)
    ctx = W.synthetic_spectra(B, 'ir', seed=1).cuda()
    nm_c, em_c = W.make_masks(n, N)
    g = torch.Generator().manual_seed(5)
    z = O.node_noise_from_raw(torch.randn(B, N, 3, generator=g), torch.randn(B, N, 6, generator=g), nm_c).cuda()
    ez = O.edge_noise_from_raw(torch.randn(B, 2, N, N, generator=g), em_c).cuda()
    outs = []
    for budget in (1 << 30, 3 * (B * N * 9 * 4 + B * 2 * N * N * 4)):
        s = _sampler(10, noise='torch')
        s.TORCH_NOISE_BUDGET = budget
        torch.manual_seed(77)
        outs.append(s.sampling(model, z, nm, em, ez, ctx))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_whole_loop_as_one_graph_launch_equals_replayed_step_graph(monkeypatch):
    """north_star: "the whole 1000-step loop is captured as a CUDA graph".  DS_LOOP_GRAPH=1 runs the loop as ONE graph launch
    (a WHILE conditional node whose body is the captured step, re-armed on the device while step < end); the default replays
    the one-step graph from the host (measured 1.6 % faster).  Same kernels, same order: bit-identical results, also for a
    second call with another segment length through the same cached graph."""
    n = torch.tensor([12, 29, 3, 17])
    B, N = 4, 29
    ctx = W.synthetic_spectra(B, 'ir', seed=1).cuda()
    outs = {}
    for flag in ('0', '1'):
        monkeypatch.setenv('DS_LOOP_GRAPH', flag)            # read by ds_create
        model = build_model('ir', salt=2, precision='bf16')
        eng = model.engine('cuda')
        plan = eng.plan(n.numpy(), N)
        emb = eng.context_embedding(ctx)
        res = []
        for steps in (7, 12):
            from diffspectra_b200.noise_schedule import ancestral_coefficients
            ns = NoiseScheduleVP('cosine', continuous_beta_0=0.1, continuous_beta_1=20.)
            coef = ancestral_coefficients(ns, torch.linspace(ns.T, 1e-3, 12, device='cuda'))
            out = eng.sample_loop(plan, emb, coef, None, None, None, seed=11, gid_base=5, steps=steps)
            res.append([t.clone() for t in out])
        outs[flag] = res
    for a, b in zip(outs['0'], outs['1']):
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert not torch.equal(outs['0'][0][0], outs['0'][1][0])          # 7 and 12 steps do differ


def test_philox_noise_matches_numpy_restatement_and_is_sharding_invariant():
    n = torch.tensor([7, 29, 2, 16])
    B, N = 4, 29
    model = build_model('ir', precision='fp32')
    eng = model.engine('cuda')
    zero, ezero = torch.zeros(B, N, 9), torch.zeros(B, N, N, 2)
    row = torch.tensor([0., 0., 1., 0.])
    seed, gid_base, step = 0x1234567890ABCDEF, 1000, 17
    plan = eng.plan(n.numpy(), N)
    ox, oe, _, _ = eng.sampler_step(plan, zero, ezero, zero, ezero, row, None, seed=seed, gid_base=gid_base, step_index=step)
    ox, oe = ox.cpu().numpy(), oe.cpu().numpy()
    for b in range(B):
        k = int(n[b])
        raw = P.node_normals(seed, gid_base + b, step, k)
        raw[:, :3] -= raw[:, :3].mean(0, keepdims=True)
        assert np.abs(ox[b, :k] - raw).max() < 5e-6
        assert np.abs(oe[b, :k, :k] - P.pair_normals(seed, gid_base + b, step, k)).max() < 5e-6
    # the same molecules in a different shard layout (gid_base shifted, batch permuted) get the same noise
    perm = [2, 0, 3, 1]
    plan2 = eng.plan(n[perm].numpy(), N)
    for j, b in enumerate(perm):
        ox2, oe2, _, _ = eng.sampler_step(plan2, zero, ezero, zero, ezero, row, None, seed=seed,
                                          gid_base=gid_base + b - j, step_index=step)
        assert np.array_equal(ox2[j].cpu().numpy(), ox[b]) and np.array_equal(oe2[j].cpu().numpy(), oe[b])
    # statistics: ~N(0,1)
    big = eng.plan(np.full(64, 29, dtype=np.int32), 29)
    z64, e64 = torch.zeros(64, 29, 9), torch.zeros(64, 29, 29, 2)
    bx, be, _, _ = eng.sampler_step(big, z64, e64, z64, e64, row, None, seed=9, gid_base=0, step_index=0)
    assert abs(bx[..., 3:].std().item() - 1) < 0.05 and abs(bx[..., 3:].mean().item()) < 0.05
    tri = be[:, torch.triu_indices(29, 29, 1)[0], torch.triu_indices(29, 29, 1)[1]]
    assert abs(tri.std().item() - 1) < 0.03 and abs(tri.mean().item()) < 0.03


def test_bf16_free_running_sampling_quality():
    """bf16 production mode on the same noise: >= 99% of atoms / bonds identical to the fp32 reference golden is
    not guaranteed for a chaotic 50-step trajectory with random weights, so only sanity is asserted here; the
    teacher-forced bf16 gate lives in test_denoiser_gpu.py."""
    from diffspectra_b200.sampling import make_masks
    g = load_golden('sampler_ir_50.pt')
    n, N, version = g['n_atoms'], g['N'], g['version']
    B = len(n)
    model = build_model(version, g['salt'], None, 'bf16')
    nm, em = make_masks(n, 'cuda', N)
    ctx = W.synthetic_spectra(B, version, seed=g['ctx_seed'])
    torch.manual_seed(g['seed'])
    nm_c, em_c = W.make_masks(n, N)
    z = O.node_noise_from_raw(torch.randn(B, N, 3), torch.randn(B, N, 6), nm_c)
    ez = O.edge_noise_from_raw(torch.randn(B, 2, N, N), em_c)
    raw = [O.draw_step_noise(B, N, nm_c, em_c) for _ in range(g['steps'])]
    rp = torch.stack([r[0] for r in raw]); rh = torch.stack([r[1] for r in raw]); re = torch.stack([r[2] for r in raw])
    eng = model.engine('cuda')
    x_mean, e_mean = eng.sample_loop(model.plan_for(nm), model.context_embedding(ctx), _sampler(50).coefficients(), z, ez,
                                     (rp, rh, re))
    print('bf16 50-step x_mean rel %.2e edge rel %.2e' % (rel_l2(x_mean, g['x_mean']), rel_l2(e_mean, g['edge_x_mean'])))
    assert torch.isfinite(x_mean).all() and torch.isfinite(e_mean).all()
    assert rel_l2(x_mean, g['x_mean']) < 5e-2 and rel_l2(e_mean, g['edge_x_mean']) < 5e-2


class _FakeMol:
    """Stand-in for one item of the reference's QM9SDataset (datasets/qm9s_dataset.py): only the attributes the eval
    driver touches (sampling.py:397-427)."""

    def __init__(self, n, seed):
        g = torch.Generator().manual_seed(seed)
        self.num_atom = torch.tensor(n)
        self.pos = torch.randn(n, 3, generator=g)
        self.rdmol = None
        self.uv = torch.log10(1 + 50 * torch.rand(1, 701, generator=g))
        self.ir = torch.log10(1 + 50 * torch.rand(1, 3501, generator=g))
        self.raman = torch.log10(1 + 50 * torch.rand(1, 3501, generator=g))


def test_eval_driver_contract_and_rank_sharding():
    """get_cond_sampling_eval_fn (sampling.py:353-468): three lists truncated to n_samples, tuples shaped like
    mol_process's; with Philox noise the union of the rank shards equals the single-rank result."""
    from diffspectra_b200.config import get_config
    from diffspectra_b200.sampling import get_cond_sampling_eval_fn
    cfg = get_config('allspectra', device='cuda', precision='bf16')
    cfg.sampling.steps = 5
    model = build_model('allspectra', salt=1, coord_scale=0.02, precision='bf16')
    ns = NoiseScheduleVP('cosine', continuous_beta_0=0.1, continuous_beta_1=20.)
    n_list = W.sample_n_atoms(40, seed=3, force_first_max=False)
    ds = [_FakeMol(int(n), 100 + i) for i, n in enumerate(n_list)]
    n_samples, batch = 21, 8

    def run(rank, world):
        fn = get_cond_sampling_eval_fn(cfg, ns, batch, n_samples, None, ds, noise='philox', seed=5, rank=rank, world_size=world)
        return fn(model)

    mols, tpos, trd = run(0, 1)
    assert len(mols) == n_samples and len(tpos) == n_samples and len(trd) == n_samples
    torch.manual_seed(42)
    perm = torch.randperm(len(ds))[:n_samples]
    for k, (pos, atom, bond, fc) in enumerate(mols):
        n = int(ds[int(perm[k])].num_atom)
        assert pos.shape == (n, 3) and atom.shape == (n,) and bond.shape == (n, n) and fc.shape == (n,)
        assert torch.equal(tpos[k], ds[int(perm[k])].pos)
        assert atom.dtype == torch.int64 and fc.dtype == torch.int64 and pos.device.type == 'cpu'
        assert torch.equal(bond, bond.t()) and bond.min() >= 0 and bond.max() <= 3 and int(atom.max()) < 5
    # two ranks: contiguous shards of the permuted list, same molecules
    shards = run(0, 2)[0] + run(1, 2)[0]
    assert len(shards) == n_samples
    same = 0
    for a, b in zip(mols, shards):
        assert a[0].shape == b[0].shape
        same += int(torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and (a[0] - b[0]).abs().max() < 1e-3)
    assert same >= n_samples - 1        # identical noise per global molecule id; bf16 tiles cut differently


def test_multi_round_driver_never_reuses_a_stale_plan():
    """Rounds of the SAME shape (B, N) but different atom counts (sampling.py:390-465): the plan / context caches of the
    model are keyed by tensor identity for speed and must never hand round r's plan to round r+2 when the allocator
    recycles the freed mask's address.  5 rounds of 8 molecules on ONE model == the same rounds each sampled alone on a
    FRESH model (Philox noise is keyed by the global sample index, so rank r of 5 is exactly round r)."""
    from diffspectra_b200.config import get_config
    from diffspectra_b200.sampling import get_cond_sampling_eval_fn, make_masks
    cfg = get_config('ir', device='cuda', precision='bf16')
    cfg.sampling.steps = 4
    ns = NoiseScheduleVP('cosine', continuous_beta_0=0.1, continuous_beta_1=20.)
    rounds, batch = 5, 8
    n_items = rounds * batch
    torch.manual_seed(42)
    perm = torch.randperm(n_items)
    n_list = W.sample_n_atoms(n_items, seed=11, force_first_max=False).clamp(max=28)
    for r in range(rounds):
        n_list[int(perm[r * batch])] = 29            # every round pads to N = 29: identical mask shapes
    ds = [_FakeMol(int(n), 500 + i) for i, n in enumerate(n_list)]
    model = build_model('ir', salt=3, coord_scale=0.02, precision='bf16')
    # (a) the cache itself: same-shape masks created and dropped in a loop (what the allocator recycles)
    for r in range(12):
        nr = n_list[perm[(r % rounds) * batch:(r % rounds + 1) * batch]]
        nm, _ = make_masks(nr.tolist(), 'cuda')
        model.engine('cuda')
        plan = model.plan_for(nm)
        assert plan.n_atoms.tolist() == nr.tolist(), (r, plan.n_atoms.tolist(), nr.tolist())
        del nm, plan
    # (b) end to end
    whole = get_cond_sampling_eval_fn(cfg, ns, batch, n_items, None, ds, noise='philox', seed=9)(model)[0]
    assert len(whole) == n_items
    for r in range(rounds):
        fresh = build_model('ir', salt=3, coord_scale=0.02, precision='bf16')
        part = get_cond_sampling_eval_fn(cfg, ns, batch, n_items, None, ds, noise='philox', seed=9, rank=r, world_size=rounds)(fresh)[0]
        assert len(part) == batch
        for k, (a, b) in enumerate(zip(whole[r * batch:(r + 1) * batch], part)):
            n_true = int(n_list[int(perm[r * batch + k])])
            assert a[0].shape == (n_true, 3) == b[0].shape
            assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3]) and torch.equal(a[0], b[0])


def test_molecule_records_kernel_equals_post_process():
    """ds_molecule_records (post_process + mol_process fused, sampling.py:12-32,53-97) == ds_post_process followed by the
    eager record packing, byte for byte; records padded to a larger rec_n unpack to the same molecules."""
    from diffspectra_b200.distributed import pack_records, record_bytes, unpack_records
    from diffspectra_b200.sampling import make_masks
    n = torch.tensor([29, 1, 7, 18, 2, 23])
    B, N = len(n), 29
    model = build_model('ir', precision='bf16')
    eng = model.engine('cuda')
    nm, em = make_masks(n, 'cuda', N)
    plan = model.plan_for(nm)
    nm_c, em_c = W.make_masks(n, N)
    g = torch.Generator().manual_seed(12)
    x = O.node_noise_from_raw(torch.randn(B, N, 3, generator=g), torch.randn(B, N, 6, generator=g) * 0.3, nm_c).cuda()
    ex = O.edge_noise_from_raw(torch.randn(B, 2, N, N, generator=g), em_c).cuda()
    pos, atom, fc, bond = eng.post_process(plan, x, ex)
    want = pack_records(pos, atom, fc, bond, torch.as_tensor(n, device='cuda'))
    got = eng.molecule_records(plan, x, ex)
    assert got.shape == (B, record_bytes(N)) and torch.equal(got, want)
    big = eng.molecule_records(plan, x, ex, rec_n=40)
    assert big.shape == (B, record_bytes(40))
    for a, b in zip(unpack_records(got, N), unpack_records(big, 40)):
        assert all(torch.equal(u, v) for u, v in zip(a, b))


def test_sharded_eval_driver_single_process():
    """diffspectra_b200.evaluate at world_size 1: same result lists as the local driver.  The
    2-GPU NCCL path is exercised by tests/run_eval_sharded.py under torchrun."""
    import subprocess
    import sys
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'):
        env.pop(k, None)
    out = subprocess.run([sys.executable, os.path.join(root, 'tests', 'run_eval_sharded.py')], env=env, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert 'EVAL_SHARDED world=1 samples=53' in out.stdout and 'OK' in out.stdout
