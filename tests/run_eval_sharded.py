"""Launched by torchrun on N >= 1 GPUs of one box (NOT collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
        tests/run_eval_sharded.py

Every rank runs diffspectra_b200.evaluate.get_cond_sampling_eval_fn (its shard + ONE NCCL all-gather of molecule
records); rank 0 also samples the whole list alone and checks that the gathered result is the same set of molecules in
the same order (Philox noise is keyed by the global sample id, so the trajectories do not depend on the sharding)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from diffspectra_b200 import evaluate as E                      # noqa: E402
from diffspectra_b200 import sampling as S                      # noqa: E402
from diffspectra_b200.config import get_config                  # noqa: E402
from diffspectra_b200.noise_schedule import NoiseScheduleVP     # noqa: E402
from oracle import weights as W                                 # noqa: E402
from tests.helpers import build_model                           # noqa: E402
from tests.test_sampler_gpu import _FakeMol                     # noqa: E402


def main():
    rank, world, dev = E.init_distributed()
    cfg = get_config('allspectra', device=str(dev), precision='bf16')
    cfg.sampling.steps = int(os.environ.get('EVAL_STEPS', '5'))
    model = build_model('allspectra', salt=1, coord_scale=0.02, precision='bf16', device=str(dev))
    ns = NoiseScheduleVP('cosine', continuous_beta_0=0.1, continuous_beta_1=20.)
    n_list = W.sample_n_atoms(90, seed=3, force_first_max=False)
    ds = [_FakeMol(int(n), 100 + i) for i, n in enumerate(n_list)]
    n_samples, batch = 53, 16
    fn = E.get_cond_sampling_eval_fn(cfg, ns, batch, n_samples, None, ds, noise='philox', seed=5)
    mols, tpos, trd = fn(model)
    assert len(mols) == n_samples == len(tpos) == len(trd), (len(mols), len(tpos))
    ok = True
    if rank == 0:
        alone = S.get_cond_sampling_eval_fn(cfg, ns, batch, n_samples, None, ds, noise='philox', seed=5, rank=0, world_size=1)(model)
        same = 0
        for a, b, p in zip(alone[0], mols, alone[1]):
            assert a[0].shape == b[0].shape
            same += int(torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and (a[0] - b[0]).abs().max() < 1e-3)
        assert all(torch.equal(x, y) for x, y in zip(alone[1], tpos))
        ok = same >= n_samples - 1
        print('EVAL_SHARDED world=%d samples=%d identical=%d %s' % (world, n_samples, same, 'OK' if ok else 'MISMATCH'), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
