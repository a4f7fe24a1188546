"""the multi-process eval driver (diffspectra_b200.evaluate) — every rank samples only its shard,
one all-gather joins them, every rank returns the reference's full result lists in the reference's order.  The
denoiser needs a GPU, so the per-rank sampler is replaced by a deterministic stand-in keyed by the global sample id;
world_size 2 over gloo must equal world_size 1."""
import os
import types

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

N_TEST, N_SAMPLES, BATCH = 23, 11, 4


class _Mol:
    def __init__(self, i):
        g = torch.Generator().manual_seed(1000 + i)
        self.num_atom = torch.tensor(int(torch.randint(1, 30, (1,), generator=g)))
        self.pos = torch.randn(int(self.num_atom), 3, generator=g)
        self.rdmol = 'rdmol-%d' % i


def _fake_mol(gid):
    g = torch.Generator().manual_seed(gid)
    n = int(torch.randint(1, 30, (1,), generator=g))
    return (torch.randn(n, 3, generator=g), torch.randint(0, 5, (n,), generator=g),
            torch.randint(0, 4, (n, n), generator=g).float(), torch.randint(-2, 3, (n,), generator=g))


def _fake_local_factory(config, noise_scheduler, batch_size, n_samples, inverse_scaler, test_ds, eps, noise, seed, rank,
                        world_size):
    """Same sharding rule as diffspectra_b200.sampling.get_cond_sampling_eval_fn, fake molecules keyed by the dataset id."""
    def fn(model):
        torch.manual_seed(42)
        perm = torch.randperm(len(test_ds))[:n_samples]
        per_rank = int(np.ceil(len(perm) / world_size))
        mine = perm[rank * per_rank:(rank + 1) * per_rank]
        return [_fake_mol(int(i)) for i in mine], None, None
    return fn


def _run(rank, world, port, q):
    from diffspectra_b200 import evaluate as E
    if world > 1:
        os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank))
        r, w, dev = E.init_distributed('gloo')
        assert (r, w) == (rank, world)
    cfg = types.SimpleNamespace(data=types.SimpleNamespace(max_node=29))
    ds = [_Mol(i) for i in range(N_TEST)]
    fn = E.get_cond_sampling_eval_fn(cfg, None, BATCH, N_SAMPLES, None, ds, _local_fn_factory=_fake_local_factory)
    model = torch.nn.Linear(1, 1)
    out = fn(model)
    if q is None:
        return out
    # numpy over the queue: torch tensors would travel as shared-memory handles that die with this process
    q.put((rank, [tuple(t.numpy().copy() for t in m) for m in out[0]], [p.numpy().copy() for p in out[1]], list(out[2])))
    dist.barrier()
    dist.destroy_process_group()


def _same(a, b):
    return len(a) == len(b) and all(torch.equal(x.float(), y.float()) for x, y in zip(a, b))


def test_sharded_eval_world2_equals_world1():
    mols1, pos1, rd1 = _run(0, 1, 0, None)
    assert len(mols1) == N_SAMPLES == len(pos1) == len(rd1)
    # ground-truth lists follow the reference's permutation (sampling.py:387-391)
    torch.manual_seed(42)
    perm = torch.randperm(N_TEST)[:N_SAMPLES]
    assert rd1 == ['rdmol-%d' % int(i) for i in perm]
    assert all(_same(m, _fake_mol(int(i))) for m, i in zip(mols1, perm))

    world, port = 2, 29741
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_run, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        r, mols, pos, rd = q.get(timeout=180)
        got[r] = ([tuple(torch.from_numpy(a) for a in m) for m in mols], [torch.from_numpy(a) for a in pos], rd)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        mols, pos, rd = got[r]
        assert rd == rd1 and _same(pos, pos1)
        assert len(mols) == N_SAMPLES                     # the odd sample count pads the last shard; padding is dropped
        for m, m1 in zip(mols, mols1):
            assert _same(m, m1)


def test_install_swaps_the_reference_seams():
    from diffspectra_b200 import evaluate as E
    from diffspectra_b200.noise_schedule import NoiseScheduleVP
    fake_run_lib = types.SimpleNamespace(get_cond_sampling_eval_fn=None, NoiseScheduleVP=None)
    E.install(fake_run_lib)
    assert fake_run_lib.NoiseScheduleVP is NoiseScheduleVP
    cfg = types.SimpleNamespace(data=types.SimpleNamespace(max_node=29, spectra_version='ir'), device='cpu', only_2D=False,
                                sampling=types.SimpleNamespace(method='ancestral', steps=5),
                                model=types.SimpleNamespace(pred_data=True, self_cond=True), pred_edge=True,
                                eval=types.SimpleNamespace(sampling_temperature=1.0))
    ns = NoiseScheduleVP('cosine', continuous_beta_0=0.1, continuous_beta_1=20.)
    fn = fake_run_lib.get_cond_sampling_eval_fn(cfg, ns, 4, 8, None, [_Mol(i) for i in range(9)])
    assert callable(fn)


def test_records_from_mols_roundtrip():
    from diffspectra_b200 import distributed as D
    from diffspectra_b200 import evaluate as E
    mols = [_fake_mol(i) for i in range(7)]
    back = D.unpack_records(E.records_from_mols(mols, 29), 29)
    for a, b in zip(mols, back):
        assert _same(a, b)
