"""CPU-side checks: parameter manifest (names / shapes / ORDER = EMA order), C-ABI symbols, host error
behaviour without a GPU, and the multi-process sharding logic (gloo, world_size 2)."""
import ctypes
import json
import os
import re

import pytest
import torch

from tests.helpers import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('kind', ['DMT', 'DMT_WO_EQ'])
@pytest.mark.parametrize('version', ['allspectra', 'ir'])
def test_parameter_manifest_matches_reference(version, kind):
    from diffspectra_b200.config import get_config
    from diffspectra_b200.model import DMT_B200, DMT_WO_EQ_B200
    man = json.load(open(os.path.join(GOLDEN, 'param_manifest.json')))[version if kind == 'DMT' else 'wo_eq_' + version]
    m = (DMT_B200 if kind == 'DMT' else DMT_WO_EQ_B200)(get_config(version, device='cpu'))
    assert [[n, list(p.shape)] for n, p in m.named_parameters()] == man['params']
    assert [[n, list(b.shape)] for n, b in m.named_buffers()] == man['buffers']
    # strict load of a DataParallel-style ('module.'-prefixed) checkpoint works on the wrapped module (utils.py:15-19)
    sd = {'module.' + k: v for k, v in m.state_dict().items()}
    torch.nn.DataParallel(m).load_state_dict(sd, strict=True)


def test_unsupported_config_is_rejected():
    from diffspectra_b200.config import get_config
    from diffspectra_b200.model import DMT_B200
    cfg = get_config('ir', device='cpu')
    cfg.model.nf = 128
    with pytest.raises(ValueError):
        DMT_B200(cfg)
    cfg = get_config('ir', device='cpu')
    cfg.data.spectra_version = 'nmr'
    with pytest.raises(ValueError):
        DMT_B200(cfg)


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as G
    G.build()
    from diffspectra_b200 import _lib as L
    header = open(os.path.join(ROOT, 'include', 'diffspectra_b200.h')).read()
    declared = set(re.findall(r'\b(ds_[a-z_0-9]+)\s*\(', header))
    assert len(declared) >= 15
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert lib.ds_version() >= 100


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_product_path_fails_loudly_without_gpu():
    from diffspectra_b200 import DiffSpectraError
    from diffspectra_b200.config import get_config
    from diffspectra_b200.model import DMT_B200
    from oracle import weights as W
    m = DMT_B200(get_config('ir', device='cpu')).eval()
    n = torch.tensor([5, 3])
    nm, em = W.make_masks(n, 5)
    x, ex = torch.zeros(2, 5, 9), torch.zeros(2, 5, 5, 2)
    with pytest.raises(DiffSpectraError), torch.no_grad():
        m(torch.zeros(2), x, nm, em, context=W.synthetic_spectra(2, 'ir'), edge_x=ex, noise_level=torch.zeros(2),
          cond_x=None, cond_edge_x=None)


def test_sampler_rejects_unsupported_modes():
    from diffspectra_b200.noise_schedule import NoiseScheduleVP
    from diffspectra_b200.sampling import AncestralSampler
    ns = NoiseScheduleVP('cosine')
    ts = torch.linspace(ns.T, 1e-3, 5)
    with pytest.raises(ValueError):
        AncestralSampler(ns, ts, False, True, True)
    with pytest.raises(ValueError):
        AncestralSampler(ns, ts, True, True, True, cond_process_fn=lambda a, b: (a.clamp(-1, 1), b))
    with pytest.raises(ValueError):
        NoiseScheduleVP('discrete')
    s = AncestralSampler(ns, ts, True, True, True, cond_process_fn=lambda a, b: (a, b))
    assert s.coefficients().shape == (5, 4)


def test_make_masks_equals_reference_loop():
    from diffspectra_b200.sampling import make_masks
    from oracle import weights as W
    n = torch.tensor([3, 29, 1, 17])
    nm, em = make_masks(n, 'cpu')
    rn, re = W.make_masks(n)
    assert torch.equal(nm, rn) and torch.equal(em, re)


def test_fma_pipe_tanh_polynomial_bound():
    """common.cuh tanh_poly2 (the FMA-pipe half of ACT_TANH_MIX): the committed coefficients, evaluated as the kernel does
    (fp32 Horner in x^2 on the clamped input), stay within 6.5e-4 of tanh on the whole real line and never leave [-1, 1]."""
    import os
    import re
    import numpy as np
    src = open(os.path.join(os.path.dirname(__file__), '..', 'diffspectra_b200', 'csrc', 'common.cuh')).read()
    c = [np.float32(re.search(r'TANH_C%d = (-?[0-9.]+e[+-][0-9]+)f' % k, src).group(1)) for k in range(9)]
    clamp = np.float32(re.search(r'fminf\(fmaxf\(x\.x, -([0-9.]+)f\)', src).group(1))
    x = np.concatenate([np.linspace(-12, 12, 600001), [-1e30, 1e30, 0.0]]).astype(np.float32)
    xc = np.clip(x, -clamp, clamp)
    t = (xc * xc).astype(np.float32)
    p = np.full_like(t, c[8])
    for k in range(7, -1, -1):
        p = (p * t + c[k]).astype(np.float32)
    p = (p * xc).astype(np.float32)
    err = np.abs(p.astype(np.float64) - np.tanh(x.astype(np.float64)))
    assert err.max() < 6.5e-4, err.max()
    assert np.abs(p).max() <= 1.0
    assert p[-1] == 0.0
