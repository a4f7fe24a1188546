"""Multi-process eval driver: keeps `main.py --mode eval` semantics (run_lib.py:297-441) with one
process per GPU instead of nn.DataParallel.

    torchrun --nnodes 1 --nproc-per-node 8 --master-addr 127.0.0.1 main.py --mode eval ...        (reference main.py)

with two lines added where the reference builds its sampling function (run_lib.py:333-337):

    from diffspectra_b200 import evaluate as b200_eval
    b200_eval.install(sys.modules[__name__])         # run_lib: registry + NoiseScheduleVP + get_cond_sampling_eval_fn

Every rank then runs the unchanged `diffspectra_evaluate`: it samples ONLY its contiguous shard of
`permute_test_mol_id[:n_samples]` (sampling.py:387-391), the shards meet in ONE all-gather of fixed-size molecule records
per `sampling_fn(model)` call (`distributed.gather_records`, NCCL over NVLink), and every rank returns the full
`(processed_mols, sampled_test_pos, sampled_test_rdkit_mols)` lists in the reference's order, so the metric code
(run_lib.py:370-441) runs unchanged (on rank 0, or redundantly on all ranks).  No collective runs inside the step loop
and no parameters are broadcast: each rank restores the checkpoint itself (utils.py:7-20).
"""
import os

import numpy as np
import torch
import torch.distributed as dist

from . import distributed as D
from . import sampling as S


def init_distributed(backend=None):
    """Process-group setup from the torchrun environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).  Returns
    (rank, world_size, device).  A single process (no RANK in the environment) needs no process group."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    cuda = torch.cuda.is_available()
    device = torch.device('cuda', local) if cuda else torch.device('cpu')
    if cuda:
        torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group(backend or ('nccl' if cuda else 'gloo'), rank=rank, world_size=world)
    return rank, world, device


def records_from_mols(mols, N):
    """mol_process tuples (pos [n,3], atom_type [n], bond [n,n], fc [n]) (sampling.py:12-32) -> [len, record_bytes(N)]
    uint8 records of `distributed.pack_records`."""
    B = len(mols)
    pos = torch.zeros(B, N, 3)
    atom = torch.zeros(B, N, dtype=torch.long)
    fc = torch.zeros(B, N, dtype=torch.long)
    bond = torch.zeros(B, N, N)
    n_atoms = torch.zeros(B, dtype=torch.long)
    for b, (p, a, e, f) in enumerate(mols):
        n = int(a.shape[0])
        if n > N:
            raise ValueError('molecule with %d atoms does not fit a record of %d' % (n, N))
        n_atoms[b] = n
        pos[b, :n] = p
        atom[b, :n] = a
        fc[b, :n] = f.reshape(-1)
        bond[b, :n, :n] = e
    return D.pack_records(pos, atom, fc, bond, n_atoms)


def gather_mols(local_mols, per_rank, N, device):
    """All ranks contribute `per_rank` records (short shards padded with empty records, n = 0) to one all-gather and
    get back the rank-major list with the padding removed."""
    rec = records_from_mols(local_mols, N)
    if rec.shape[0] < per_rank:
        rec = torch.cat([rec, torch.zeros(per_rank - rec.shape[0], rec.shape[1], dtype=torch.uint8)])
    use_cuda = dist.is_initialized() and dist.get_backend() == 'nccl'
    rec = D.gather_records(rec.to(device if use_cuda else 'cpu'))
    mols = D.unpack_records(rec, N)
    return [m for m in mols if m[1].shape[0] > 0]


def get_cond_sampling_eval_fn(config, noise_scheduler, batch_size, n_samples, inverse_scaler, test_ds, eps=1e-3,
                              noise='philox', seed=0, _local_fn_factory=None):
    """Reference signature (sampling.py:353); `sampling_fn(model)` returns the FULL result lists on every rank.
    noise='philox' (default here) keys the device RNG by the global sample index, so the generated molecules do not
    depend on the number of GPUs; noise='torch' reproduces the reference's generator stream on one GPU only."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world > 1 and noise != 'philox':
        raise ValueError("multi-process eval needs noise='philox' (the torch generator stream is not shardable)")
    factory = _local_fn_factory or S.get_cond_sampling_eval_fn
    local_fn = factory(config, noise_scheduler, batch_size, n_samples, inverse_scaler, test_ds, eps, noise=noise, seed=seed,
                       rank=rank, world_size=world)
    N = int(getattr(config.data, 'max_node', 0) or 29)

    def sampling_fn(model):
        total = min(n_samples, len(test_ds))
        per_rank = int(np.ceil(total / world))
        if world > 1 and hasattr(local_fn, 'local_records') and dist.get_backend() == 'nccl':
            # product path: the shard's records never leave the device before the all-gather; ONE D2H copy afterwards
            rec, rn, _, _ = local_fn.local_records(model)
            dev = next(model.parameters()).device
            if rec is None:
                rec = torch.zeros(0, D.record_bytes(rn), dtype=torch.uint8, device=dev)
            if rec.shape[0] < per_rank:                 # short last shard: empty records (n = 0) keep the gather uniform
                rec = torch.cat([rec, torch.zeros(per_rank - rec.shape[0], rec.shape[1], dtype=torch.uint8, device=dev)])
            mols = [m for m in D.unpack_records(D.gather_records(rec), rn) if m[1].shape[0] > 0]
        else:
            local_mols, _, _ = local_fn(model)
            if world == 1:
                mols = local_mols
            else:
                dev = next(model.parameters()).device
                mols = gather_mols(local_mols, per_rank, N, dev)
        # ground truth of ALL samples, same permutation as the local driver (sampling.py:387-391)
        g = torch.Generator().manual_seed(42)
        perm = torch.randperm(len(test_ds), generator=g)[:n_samples]
        _, _, test_pos, test_rdmols = S.stage_round(test_ds, perm, [])      # batched for an in-memory data set
        return mols[:n_samples], test_pos[:n_samples], test_rdmols[:n_samples]

    return sampling_fn


def install(run_lib_module, noise='philox', seed=0, precision=None):
    """Swap the three hot-path symbols of the reference's `run_lib` namespace for the B200 ones and register the models:
    `create_model` then resolves `--config.model.name DMT_B200 | DMT_WO_EQ_B200` (models/utils.py:5-28), and
    `diffspectra_evaluate` (run_lib.py:297-441) runs unchanged, one process per GPU."""
    from . import model as M
    from .noise_schedule import NoiseScheduleVP
    try:
        M.register()
    except ImportError:            # reference `models` package not importable: nothing to register into
        pass
    init_distributed()

    def _factory(config, noise_scheduler, batch_size, n_samples, inverse_scaler, test_ds, eps=1e-3):
        if precision is not None:
            config.model.b200_precision = precision
        return get_cond_sampling_eval_fn(config, noise_scheduler, batch_size, n_samples, inverse_scaler, test_ds, eps,
                                         noise=noise, seed=seed)
    run_lib_module.get_cond_sampling_eval_fn = _factory
    run_lib_module.NoiseScheduleVP = NoiseScheduleVP
    return run_lib_module
