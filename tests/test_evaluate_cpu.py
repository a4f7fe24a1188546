"""the multi-process eval driver (diffspectra_b200.evaluate) — every rank samples only its shard,
one all-gather joins them, every rank returns the reference's full result lists in the reference's order.  The
denoiser needs a GPU, so the per-rank sampler is replaced by a deterministic stand-in keyed by the global sample id;
world_size 2 over gloo must equal world_size 1."""
import os
import types

import pytest

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
    from diffspectra_b200.config import get_config
    cfg = get_config('ir', device='cpu')
    cfg.sampling.steps = 5
    ns = NoiseScheduleVP('cosine', continuous_beta_0=0.1, continuous_beta_1=20.)
    fn = fake_run_lib.get_cond_sampling_eval_fn(cfg, ns, 4, 8, None, [_Mol(i) for i in range(9)])
    assert callable(fn)
    # a data configuration the record / post-process kernels do not implement must be refused, not mis-sampled
    for key, section, bad in (('normalize_factors', 'model', '1, 2, 2, 1'), ('compress_edge', 'data', False),
                              ('centered', 'data', False), ('atom_types', 'data', 4), ('include_fc_charge', 'model', False)):
        cfg2 = get_config('ir', device='cpu')
        cfg2[section][key] = bad
        with pytest.raises(ValueError):
            fake_run_lib.get_cond_sampling_eval_fn(cfg2, ns, 4, 8, None, [_Mol(i) for i in range(9)])


def test_stage_round_fast_path_equals_item_loop():
    """sampling.stage_round: the InMemoryDataset fast path (one index_select per spectrum over the collated storage)
    returns exactly what the reference's item-by-item loop (sampling.py:397-427) builds."""
    from diffspectra_b200.sampling import stage_round
    g = torch.Generator().manual_seed(0)
    n_items = 12
    n_at = torch.randint(3, 10, (n_items,), generator=g)
    off = torch.cat([torch.zeros(1, dtype=torch.long), n_at.cumsum(0)])

    class Item:
        pass

    class InMem:                                   # the two attributes PyG's InMemoryDataset keeps: _data + slices
        def __init__(self, indices=None):
            self._data = types.SimpleNamespace(
                uv=torch.rand(n_items, 701, generator=torch.Generator().manual_seed(1)),
                ir=torch.rand(n_items, 3501, generator=torch.Generator().manual_seed(2)),
                raman=torch.rand(n_items, 3501, generator=torch.Generator().manual_seed(3)),
                num_atom=n_at.clone(), pos=torch.rand(int(off[-1]), 3, generator=torch.Generator().manual_seed(4)))
            ar = torch.arange(n_items + 1)
            self.slices = dict(uv=ar, ir=ar, raman=ar, num_atom=ar, pos=off)
            self._indices = indices

        def __len__(self):
            return n_items if self._indices is None else len(self._indices)

        def __getitem__(self, i):
            j = int(i) if self._indices is None else int(self._indices[int(i)])
            it = Item()
            it.uv, it.ir, it.raman = self._data.uv[j:j + 1], self._data.ir[j:j + 1], self._data.raman[j:j + 1]
            it.num_atom = self._data.num_atom[j]
            it.pos = self._data.pos[int(off[j]):int(off[j + 1])]
            it.rdmol = None
            return it

    for ds in (InMem(), InMem(indices=[7, 2, 9, 0, 11, 4])):
        ids = [3, 0, 5] if len(ds) < 12 else [10, 1, 7, 3]
        keys = ['uv', 'ir', 'raman']
        fast = stage_round(ds, ids, keys)
        items = [ds[i] for i in ids]
        assert fast[0] == [int(m.num_atom) for m in items]
        for k, t in zip(keys, fast[1]):
            assert t.shape == (len(ids), 1, getattr(items[0], k).shape[-1])
            assert torch.equal(t, torch.stack([getattr(m, k) for m in items]))
        assert all(torch.equal(a, m.pos) for a, m in zip(fast[2], items))
    # a plain list takes the item loop
    plain = [InMem()[i] for i in range(5)]
    out = stage_round(plain, [4, 2], ['ir'])
    assert out[0] == [int(plain[4].num_atom), int(plain[2].num_atom)] and out[1][0].shape == (2, 1, 3501)


def test_records_from_mols_roundtrip():
    from diffspectra_b200 import distributed as D
    from diffspectra_b200 import evaluate as E
    mols = [_fake_mol(i) for i in range(7)]
    back = D.unpack_records(E.records_from_mols(mols, 29), 29)
    for a, b in zip(mols, back):
        assert _same(a, b)
