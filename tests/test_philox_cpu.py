"""The numpy restatement of the device noise source (oracle/philox_ref.py) against the PUBLISHED Philox4x32-10 known-answer
vectors (Random123 `kat_vectors`, D. E. Shaw Research), plus the properties the sampler relies on.  The GPU tests then
compare the CUDA kernel with this restatement bit for bit at the uniform level (tests/test_sampler_gpu.py)."""
import numpy as np

from oracle import philox_ref as P

KAT = [  # (counter, key, philox4x32-10 output)
    ((0x00000000,) * 4, (0x00000000, 0x00000000), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox4x32_10_known_answers():
    for ctr, key, exp in KAT:
        out = P.philox4x32_10(np.array([ctr], dtype=np.uint32), key)[0]
        assert tuple(int(v) for v in out) == exp
    # vectorised call == element-wise calls
    ctrs = np.array([k[0] for k in KAT], dtype=np.uint32)
    same_key = KAT[2][1]
    batch = P.philox4x32_10(ctrs, same_key)
    for i in range(len(KAT)):
        assert np.array_equal(batch[i], P.philox4x32_10(ctrs[i:i + 1], same_key)[0])


def test_noise_is_keyed_by_global_id_and_step_only():
    """What makes the sampling loop invariant to sharding: the draw of a molecule depends on
    (seed, global molecule id, step, atom / pair index) and on nothing else — not on the batch or the rank."""
    a = P.node_normals(seed=7, gid=123456, step=5, n_atoms=11)
    b = P.node_normals(seed=7, gid=123456, step=5, n_atoms=29)[:11]
    assert np.array_equal(a, b)                                   # independent of the molecule's size / padding
    assert not np.array_equal(a, P.node_normals(7, 123457, 5, 11))
    assert not np.array_equal(a, P.node_normals(7, 123456, 6, 11))
    assert not np.array_equal(a, P.node_normals(8, 123456, 5, 11))
    e = P.pair_normals(seed=7, gid=123456, step=5, n_atoms=9)
    assert np.array_equal(e, e.transpose(1, 0, 2))                # symmetric edge noise (models/utils.py:100-106)
    assert np.all(e[np.arange(9), np.arange(9)] == 0)
    assert np.array_equal(e, P.pair_normals(7, 123456, 5, 29)[:9, :9])
    # 64-bit ids: the high word enters the counter
    assert not np.array_equal(P.node_normals(7, 5, 0, 4), P.node_normals(7, 5 + (1 << 32), 0, 4))


def test_normals_are_standard_normal():
    x = np.concatenate([P.node_normals(3, g, s, 29).ravel() for g in range(40) for s in range(5)])
    assert abs(x.mean()) < 0.02 and abs(x.std() - 1.0) < 0.02
    assert abs(np.mean(x ** 3)) < 0.06 and abs(np.mean(x ** 4) - 3.0) < 0.15
    assert np.isfinite(x).all() and np.abs(x).max() < 6.0
