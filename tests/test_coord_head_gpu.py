"""The fused coordinate head (csrc/coord_head_tc.cu; models/dmt.py:37-60 MultiCondEquiUpdate) on CTA pairs.

* ds_umma2_probe pins the operand / accumulator split of one cta_group::2 tcgen05.mma (M = 256 over two CTAs).
* ds_coord_head on synthetic per-atom / per-pair operands against a torch fp32 restatement of the same maths, for batches
  from one molecule to the configs[1] size (every CTA pair walks many tiles; windows, tails and empty molecules)."""
import ctypes

import numpy as np
import pytest
import torch

from diffspectra_b200 import _lib as L
from oracle import weights as W

pytestmark = pytest.mark.gpu
ADA_LD, ADA_COORD = 19584, 1920


@pytest.fixture(scope='module')
def ctx():
    h = ctypes.c_void_p()
    L.check(L.lib().ds_create(ctypes.byref(h), 0, L.MODE_BF16, 3), 'ds_create')
    yield h
    L.lib().ds_destroy(h)


@pytest.mark.parametrize('K', [64, 128, 256])
def test_cta_pair_mma_probe(ctx, K):
    g = torch.Generator(device='cuda').manual_seed(K)
    A = torch.randn(256, K, device='cuda', generator=g).bfloat16()
    Wt = (torch.randn(256, K, device='cuda', generator=g) / K ** 0.5).bfloat16()
    out = torch.full((256, 256), float('nan'), device='cuda')
    L.check(L.lib().ds_umma2_probe(ctx, L.ptr(A), L.ptr(Wt), L.ptr(out), K, L.stream_ptr()), 'ds_umma2_probe')
    torch.cuda.synchronize()
    ref = A.float() @ Wt.float().t()
    err = (out - ref).abs()
    blocks = [[err[128 * i:128 * i + 128, 128 * j:128 * j + 128].max().item() for j in range(2)] for i in range(2)]
    print('cta-pair MMA K=%d max err per 128x128 block' % K, blocks)
    assert torch.isfinite(out).all() and err.max().item() < 2e-3


def _plan_tables(n_atoms):
    """pairs (row-major upper triangle per molecule) and the source-major directed-edge index, as in the plan."""
    noff = np.concatenate([[0], np.cumsum(n_atoms)])
    poff = np.concatenate([[0], np.cumsum(n_atoms * (n_atoms - 1) // 2)])
    mol, ri, rj, dfwd, drev = [], [], [], [], []
    for b, n in enumerate(n_atoms):
        iu = np.triu_indices(n, 1)
        mol.append(np.full(len(iu[0]), b))
        ri.append(noff[b] + iu[0])
        rj.append(noff[b] + iu[1])
        dfwd.append(2 * poff[b] + iu[0] * (n - 1) + iu[1] - 1)
        drev.append(2 * poff[b] + iu[1] * (n - 1) + iu[0])
    cat = lambda v: torch.from_numpy(np.concatenate(v)).long().cuda()
    return cat(mol), cat(ri), cat(rj), cat(dfwd), cat(drev), int(noff[-1]), int(poff[-1])


@pytest.mark.parametrize('case', ['one', 'small', 'tiny_molecules', 'configs1', 'n64'])
def test_coord_head_matches_torch(ctx, case):
    from diffspectra_b200.engine import Plan
    N = 29
    if case == 'one':
        n = np.array([18], dtype=np.int32)
    elif case == 'small':
        n = np.array([29, 1, 4, 17, 2, 23, 9, 1, 1, 12, 3, 29], dtype=np.int32)
    elif case == 'tiny_molecules':           # many molecules per 128-pair tile: atoms beyond the 48-atom window
        n = np.tile(np.array([3, 2, 4, 1, 5], dtype=np.int32), 60)
    elif case == 'configs1':
        n = W.sample_n_atoms(1024, seed=1234).numpy().astype(np.int32)
    else:
        N = 64
        n = np.array([64, 40, 64, 33, 64, 1, 50], dtype=np.int32)

    class _Eng:                               # Plan only needs the handle and the device
        h, device = ctx, torch.device('cuda')
    plan = Plan(_Eng, n, N)
    mol, ri, rj, dfwd, drev, Mn, Mp = _plan_tables(n)
    assert (Mn, Mp) == (plan.Mn, plan.Mp)
    B = len(n)
    g = torch.Generator(device='cuda').manual_seed(len(n) + N)
    X = torch.randn(Mp, 128, device='cuda', generator=g).bfloat16()
    ab = (torch.randn(Mn, 512, device='cuda', generator=g) * 0.7).bfloat16()
    we = (torch.randn(256, 128, device='cuda', generator=g) / 128 ** 0.5).bfloat16()
    wc1 = torch.randn(256, 256, device='cuda', generator=g) / 16
    bc1 = torch.randn(256, device='cuda', generator=g) * 0.1
    wc1h, bc1h = (wc1 * 0.5).bfloat16(), bc1 * 0.5              # first layer pre-halved: SiLU(2 (z Wh^T + bh))
    wc2 = torch.randn(3, 256, device='cuda', generator=g) / 16
    ada = torch.zeros(B, ADA_LD, device='cuda')
    ada[:, ADA_COORD:ADA_COORD + 512] = torch.randn(B, 512, device='cuda', generator=g) * 0.3
    pflags = torch.randint(0, 4, (max(Mp, 1),), device='cuda', generator=g, dtype=torch.uint8)
    wdir = torch.full((2 * Mp + 8,), -7.0, device='cuda')
    scratch = torch.empty(B * 1024, dtype=torch.uint8, device='cuda')
    L.check(L.lib().ds_coord_head(ctx, *plan.args(), L.ptr(X), L.ptr(ab), L.ptr(ada), L.ptr(pflags), L.ptr(we), L.ptr(wc1h),
                                  L.ptr(bc1h), L.ptr(wc2), L.ptr(wdir), L.ptr(scratch), L.stream_ptr()), 'ds_coord_head')
    torch.cuda.synchronize()
    # torch restatement (fp32 on the same bf16 operands; the kernel rounds z to bf16 before coord_mlp.0)
    G = X.float() @ we.float().t()
    A_, B_ = ab[:, :256].float(), ab[:, 256:].float()
    sh, sc = ada[mol, ADA_COORD:ADA_COORD + 256], ada[mol, ADA_COORD + 256:ADA_COORD + 512]
    adj = torch.stack([torch.ones(Mp, device='cuda'), (pflags[:Mp] & 1).float(), ((pflags[:Mp] >> 1) & 1).float()], dim=1)
    ref = torch.full_like(wdir, -7.0)
    for src, dst, d in ((ri, rj, dfwd), (rj, ri, drev)):
        y = A_[src] + B_[dst] + G
        z = torch.nn.functional.layer_norm(y, (256,), eps=1e-6) * (1 + sc) + sh
        u = torch.nn.functional.silu(z.bfloat16().float() @ (wc1h.float() * 2).t() + bc1)
        ref[d] = (torch.tanh(u @ wc2.t()) * adj).mean(-1)
    assert sorted(torch.cat([dfwd, drev]).tolist()) == list(range(2 * Mp))
    err = (wdir[:2 * Mp] - ref[:2 * Mp]).abs()
    print(case, 'Mn %d Mp %d max err %.2e mean err %.2e' % (Mn, Mp, err.max().item() if Mp else 0, err.mean().item() if Mp else 0))
    assert torch.isfinite(wdir).all()
    # bf16 roundings inside the kernel the fp32 restatement does not have: A + B, y, LN(y), the modulate vectors, the
    # coord_mlp.2 / bias table; |w| <= 1
    assert err.max().item() < 1.5e-2 and err.mean().item() < 2e-3
    assert (wdir[2 * Mp:] == -7).all()
