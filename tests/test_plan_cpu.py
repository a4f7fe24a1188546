"""Host logic of the molecule plan (ds_plan_build): every table of the packed ragged layout against an independent numpy
restatement.  Runs without a GPU through ds_plan_build_host, which fills the very blob ds_plan_build copies to the device.
Reference counterparts: adj_mask.nonzero() + dense_to_sparse() row-major (b, i, j) order (models/dmt.py:327-329)."""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ['n_atoms', 'noff', 'poff', 'node_info', 'pair_info', 'dir_info', 'dir_mol', 'pair_rows', 'mol_order', 'node_order',
         'mol_launch', 'atom_launch', 'total']


def _lib():
    lib = ctypes.CDLL(os.path.join(ROOT, 'diffspectra_b200', 'libdiffspectra_b200.so'))
    lib.ds_plan_bytes.restype = ctypes.c_size_t
    lib.ds_last_error.restype = ctypes.c_char_p
    return lib


def _build(lib, n_atoms, N):
    n_atoms = np.ascontiguousarray(n_atoms, dtype=np.int32)
    B = len(n_atoms)
    off = (ctypes.c_size_t * 13)()
    assert lib.ds_plan_layout(B, N, off, 13) == 0
    off = dict(zip(NAMES, list(off)))
    nbytes = lib.ds_plan_bytes(B, N)
    assert nbytes == off['total']
    buf = np.zeros(nbytes, dtype=np.uint8)
    mn, mp = ctypes.c_int(), ctypes.c_int()
    rc = lib.ds_plan_build_host(n_atoms.ctypes.data_as(ctypes.c_void_p), B, N, buf.ctypes.data_as(ctypes.c_void_p),
                                ctypes.c_size_t(nbytes), ctypes.byref(mn), ctypes.byref(mp))
    return rc, buf, off, mn.value, mp.value


def _view(buf, off, name, count, dtype=np.int32, width=1):
    a = buf[off[name]:off[name] + count * width * 4].view(dtype)
    return a.reshape(count, width) if width > 1 else a


@pytest.mark.parametrize('N,seed', [(29, 0), (29, 1), (64, 2), (5, 3)])
def test_plan_tables_match_numpy_restatement(N, seed):
    lib = _lib()
    rng = np.random.default_rng(seed)
    B = 37
    n = rng.integers(1, N + 1, size=B).astype(np.int32)
    n[0], n[1] = N, 1                                     # a full molecule and a single atom (no pairs)
    rc, buf, off, Mn, Mp = _build(lib, n, N)
    assert rc == 0, lib.ds_last_error()
    assert Mn == int(n.sum()) and Mp == int((n * (n - 1) // 2).sum())
    noff = np.concatenate([[0], np.cumsum(n)]).astype(np.int32)
    poff = np.concatenate([[0], np.cumsum(n * (n - 1) // 2)]).astype(np.int32)
    assert np.array_equal(_view(buf, off, 'n_atoms', B), n)
    assert np.array_equal(_view(buf, off, 'noff', B + 1), noff)
    assert np.array_equal(_view(buf, off, 'poff', B + 1), poff)
    node_info, pair_info, pair_rows, dir_info = [], [], [], []
    for b in range(B):
        nb = int(n[b])
        node_info += [(b << 6) | i for i in range(nb)]
        loc = {}
        for i in range(nb):
            for j in range(i + 1, nb):                    # row-major upper triangle
                loc[(i, j)] = poff[b] + len(loc)
                pair_info.append((b << 12) | (i << 6) | j)
                pair_rows.append((noff[b] + i, noff[b] + j))
        for r in range(nb):                               # directed edges, source-major, the reference's (b, i, j) order
            for c in range(nb):
                if c != r:
                    dir_info.append((loc[(min(r, c), max(r, c))], noff[b] + r, noff[b] + c, b))
    assert np.array_equal(_view(buf, off, 'node_info', Mn, np.uint32), np.array(node_info, dtype=np.uint32))
    assert np.array_equal(_view(buf, off, 'pair_info', Mp, np.uint32), np.array(pair_info, dtype=np.uint32))
    assert np.array_equal(_view(buf, off, 'pair_rows', Mp, np.int32, 2), np.array(pair_rows, dtype=np.int32).reshape(Mp, 2))
    di = np.array(dir_info, dtype=np.int32).reshape(2 * Mp, 4)
    assert np.array_equal(_view(buf, off, 'dir_info', 2 * Mp, np.int32, 4), di)
    assert np.array_equal(_view(buf, off, 'dir_mol', 2 * Mp, np.uint32), di[:, 3].astype(np.uint32))
    # launch order: a permutation of the molecules, atom counts non-increasing, stable
    order = _view(buf, off, 'mol_order', B)
    assert np.array_equal(order, np.argsort(-n, kind='stable'))
    ml = _view(buf, off, 'mol_launch', B, np.int32, 4)
    assert np.array_equal(ml, np.stack([order, n[order], noff[order], poff[order]], axis=1))
    node_order = _view(buf, off, 'node_order', Mn)
    al = _view(buf, off, 'atom_launch', Mn, np.int32, 4)
    exp_nodes = np.concatenate([noff[m] + np.arange(n[m]) for m in order]).astype(np.int32)
    assert np.array_equal(node_order, exp_nodes) and sorted(node_order.tolist()) == list(range(Mn))
    exp_al = np.concatenate([np.stack([noff[m] + np.arange(n[m]), np.full(n[m], m), (n[m] << 8) | np.arange(n[m]),
                                       np.full(n[m], poff[m])], axis=1) for m in order]).astype(np.int32)
    assert np.array_equal(al, exp_al)


def test_plan_rejects_bad_arguments():
    lib = _lib()
    for n, N in (([0, 3], 5), ([6, 3], 5), ([3], 65), ([3], 0)):
        if N <= 0:
            assert lib.ds_plan_bytes(1, N) == 0
            continue
        rc, *_ = _build(lib, np.array(n), N) if N <= 64 else (None,)
        if N > 64:
            buf = np.zeros(1 << 20, dtype=np.uint8)
            mn, mp = ctypes.c_int(), ctypes.c_int()
            arr = np.array(n, dtype=np.int32)
            rc = lib.ds_plan_build_host(arr.ctypes.data_as(ctypes.c_void_p), 1, N, buf.ctypes.data_as(ctypes.c_void_p),
                                        ctypes.c_size_t(buf.size), ctypes.byref(mn), ctypes.byref(mp))
        assert rc != 0 and b'ds_plan_build' in lib.ds_last_error()
