"""Name-keyed deterministic weights + synthetic inputs shared by the reference (build container)
and by our module (GPU box).  TEST INFRASTRUCTURE ONLY.

Full-width DMT weights are 130-160 MB; instead of shipping them, every tensor of a
reference-compatible ``state_dict`` is filled from a generator seeded by crc32(name), so the
reference module here and ``DMT_B200`` on the GPU box hold bit-identical parameters
(SURVEY.md §8(c)).  Distributions follow the reference's own initialisers in spirit (uniform
+-1/sqrt(fan_in) for Linear; U(0,3) for RBF means/stds, models/layers.py:325-326; U(-0.02,0.02)
for positional tables, specformer_layers.py:105-107) but biases / BatchNorm statistics are made
non-trivial so that every term of the maths is exercised.
"""
import zlib

import torch

# the synthetic-input helpers live in the product package (bench.py's CUDA arm may not import oracle/); re-exported here
# because every test and the golden generator reach them as W.sample_n_atoms / W.synthetic_spectra
from diffspectra_b200.synthetic import QM9_N_NODES, sample_n_atoms, synthetic_spectra  # noqa: E402,F401


def _gen(name, salt):
    g = torch.Generator()
    g.manual_seed((zlib.crc32(name.encode()) + 7919 * salt) & 0x7FFFFFFF)
    return g


def keyed_fill_(state_dict, salt=0, coord_scale=None):
    """In-place deterministic fill of every floating tensor of a DMT-compatible state_dict."""
    for name, t in state_dict.items():
        if not torch.is_floating_point(t):
            continue                                    # num_batches_tracked
        g = _gen(name, salt)
        shape = tuple(t.shape)

        def U(lo, hi):
            return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo

        if name.endswith('sdp_attn.scale'):
            v = torch.tensor(8 ** -0.5)
        elif name.endswith('coord_norm.scale'):
            v = torch.full(shape, 1e-2 if coord_scale is None else float(coord_scale))
        elif name.endswith('means.weight') or name.endswith('stds.weight'):
            v = U(0., 3.)
        elif 'W_pos' in name:
            v = U(-0.02, 0.02)
        elif name.endswith('time_mlp.0.weights'):
            v = torch.randn(shape, generator=g, dtype=torch.float32)
        elif name.endswith('running_mean'):
            v = U(-0.2, 0.2)
        elif name.endswith('running_var'):
            v = U(0.5, 1.5)
        elif ('norm_attn' in name or 'norm_ffn' in name or 'out_norm' in name) and name.endswith('weight'):
            v = U(0.8, 1.2)
        elif ('norm_attn' in name or 'norm_ffn' in name or 'out_norm' in name) and name.endswith('bias'):
            v = U(-0.1, 0.1)
        elif name.endswith('.weight') and t.dim() == 2:
            bound = 1.0 / (shape[1] ** 0.5)
            v = U(-bound, bound)
        elif name.endswith('.bias'):
            v = U(-0.05, 0.05)
        else:
            raise KeyError('keyed_fill_: no rule for %s %s' % (name, shape))
        t.copy_(v.to(t.dtype).reshape(shape))
    return state_dict


def make_masks(n_atoms, N=None):
    """node_mask [B,N,1], edge_mask [B*N*N,1] exactly as sampling.py:432-439."""
    B = len(n_atoms)
    N = int(max(n_atoms)) if N is None else N
    node_mask = torch.zeros(B, N)
    for i in range(B):
        node_mask[i, 0:int(n_atoms[i])] = 1
    edge_mask = node_mask.unsqueeze(1) * node_mask.unsqueeze(2)
    diag = ~torch.eye(N, dtype=torch.bool).unsqueeze(0)
    edge_mask = edge_mask * diag
    return node_mask.unsqueeze(2), edge_mask.view(B * N * N, 1)
