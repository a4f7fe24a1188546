"""numpy restatement of the device noise generator of sampler_kernels.cu (Philox4x32-10 + Box-Muller) used by the
throughput path; lets the tests check the in-kernel noise bit-for-bit at the uniform level and to ~1e-6 after the
transcendental functions.  TEST INFRASTRUCTURE ONLY."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
TAG_NODE, TAG_EDGE = 0x4e4f4445, 0x45444745


def philox4x32_10(c, k):
    """c: [..., 4] uint32 counters, k: (k0, k1) uint32."""
    c = [c[..., i].astype(np.uint32) for i in range(4)]
    k0, k1 = np.uint32(k[0]), np.uint32(k[1])
    for _ in range(10):
        p0 = M0 * c[0].astype(np.uint64)
        p1 = M1 * c[2].astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        with np.errstate(over='ignore'):
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return np.stack(c, axis=-1)


def box_muller(a, b):
    u1 = ((a >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    u2 = ((b >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    r = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    th = (np.float32(6.28318530717958647692) * u2).astype(np.float32)
    return (r * np.cos(th)).astype(np.float32), (r * np.sin(th)).astype(np.float32)


def node_normals(seed, gid, step, n_atoms):
    """[n_atoms, 9] raw normals of one molecule (step = -1 for the initial draw)."""
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.zeros((n_atoms, 12), dtype=np.float32)
    for q in range(3):
        ctr = np.zeros((n_atoms, 4), dtype=np.uint32)
        ctr[:, 0] = np.arange(n_atoms) + 64 * q
        ctr[:, 1] = np.uint32(step & 0xFFFFFFFF)
        ctr[:, 2] = np.uint32(gid & 0xFFFFFFFF)
        ctr[:, 3] = np.uint32(TAG_NODE ^ ((gid >> 32) & 0xFFFFFFFF))
        r = philox4x32_10(ctr, key)
        n0x, n0y = box_muller(r[:, 0], r[:, 1])
        n1x, n1y = box_muller(r[:, 2], r[:, 3])
        out[:, q * 4 + 0], out[:, q * 4 + 1], out[:, q * 4 + 2], out[:, q * 4 + 3] = n0x, n0y, n1x, n1y
    return out[:, :9]


def pair_normals(seed, gid, step, n_atoms):
    """[n, n, 2] symmetric raw normals of one molecule (zero diagonal)."""
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.zeros((n_atoms, n_atoms, 2), dtype=np.float32)
    ii, jj = np.triu_indices(n_atoms, 1)
    if len(ii) == 0:
        return out
    ctr = np.zeros((len(ii), 4), dtype=np.uint32)
    ctr[:, 0] = ii * 64 + jj
    ctr[:, 1] = np.uint32(step & 0xFFFFFFFF)
    ctr[:, 2] = np.uint32(gid & 0xFFFFFFFF)
    ctr[:, 3] = np.uint32(TAG_EDGE ^ ((gid >> 32) & 0xFFFFFFFF))
    r = philox4x32_10(ctr, key)
    z0, z1 = box_muller(r[:, 0], r[:, 1])
    out[ii, jj, 0], out[ii, jj, 1] = z0, z1
    out[jj, ii, 0], out[jj, ii, 1] = z0, z1
    return out
