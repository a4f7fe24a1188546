"""Synthetic QM9S-shaped inputs for benchmarks, smoke runs and tests: atom counts from the QM9S histogram, spectra of
the data set's shapes, and an in-memory data set with the interface the eval driver consumes.  No chemistry: there is
no network for the real data (BASELINE.json: "measured on synthetic molecules and spectra of the QM9S config's shapes").
"""
import types

import torch

# atom-count histogram of QM9S (datasets/datasets_config.py:23-25, 'qm9_second_half')
QM9_N_NODES = {3: 1, 4: 3, 5: 3, 6: 5, 7: 7, 8: 25, 9: 62, 10: 178, 11: 412, 12: 845, 13: 1541, 14: 2587,
               15: 3865, 16: 5344, 17: 6461, 18: 6695, 19: 6944, 20: 4794, 21: 4962, 22: 1701, 23: 2380,
               24: 267, 25: 754, 26: 17, 27: 132, 29: 15}
SPECTRUM_LEN = {'uv': 701, 'ir': 3501, 'raman': 3501}


def sample_n_atoms(B, seed=1234, force_first_max=True, max_n=29):
    """n ~ Categorical(QM9S histogram); molecule 0 forced to max_n so N_pad = 29."""
    ks = torch.tensor(sorted(QM9_N_NODES.keys()))
    w = torch.tensor([QM9_N_NODES[int(k)] for k in ks], dtype=torch.float64)
    g = torch.Generator()
    g.manual_seed(seed)
    idx = torch.multinomial(w / w.sum(), B, replacement=True, generator=g)
    n = ks[idx].clone()
    if force_first_max:
        n[0] = max_n
    return n


def synthetic_spectra(B, version='allspectra', seed=1235):
    """log10(1 + 50*U[0,1)) spectra (mirrors datasets/build_dataset.py:142-148); list or tensor like
    sampling.py:423-427."""
    g = torch.Generator()
    g.manual_seed(seed)
    order = ['uv', 'ir', 'raman'] if version == 'allspectra' else [version]
    out = [torch.log10(1 + 50 * torch.rand(B, 1, SPECTRUM_LEN[k], generator=g)) for k in order]
    return out if version == 'allspectra' else out[0]


class SyntheticQM9S:
    """In-memory stand-in for the reference's test split (datasets/qm9s_dataset.py: a PyG InMemoryDataset): items with
    `.num_atom`, `.pos`, `.rdmol`, `.uv`/`.ir`/`.raman` ([1, L]); the collated storage is exposed as `_data` + `slices`
    like PyG does, so `sampling.stage_round` takes its batched path.  Spectra live in (optionally pinned) host memory."""

    def __init__(self, n_items, version='allspectra', seed=1234, n_atoms=None, pin=False):
        n = sample_n_atoms(n_items, seed=seed, force_first_max=False) if n_atoms is None else torch.as_tensor(n_atoms).long()
        self.keys = ['uv', 'ir', 'raman'] if version == 'allspectra' else [version]
        g = torch.Generator().manual_seed(seed + 1)
        d = types.SimpleNamespace()
        for k in self.keys:
            t = torch.log10(1 + 50 * torch.rand(n_items, SPECTRUM_LEN[k], generator=g))
            setattr(d, k, t.pin_memory() if pin else t)
        d.num_atom = n.clone()
        off = torch.cat([torch.zeros(1, dtype=torch.long), n.cumsum(0)])
        d.pos = torch.randn(int(off[-1]), 3, generator=g)
        ar = torch.arange(n_items + 1)
        self._data = d
        self.slices = {k: ar for k in self.keys}
        self.slices.update(num_atom=ar, pos=off)
        self._indices = None
        self._off = off

    def __len__(self):
        return int(self._data.num_atom.shape[0])

    def __getitem__(self, i):
        i = int(i)
        it = types.SimpleNamespace(num_atom=self._data.num_atom[i], rdmol=None,
                                   pos=self._data.pos[int(self._off[i]):int(self._off[i + 1])])
        for k in self.keys:
            setattr(it, k, getattr(self._data, k)[i:i + 1])
        return it
