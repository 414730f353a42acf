"""Shared helpers for the test-suite (golden loading, seeded inputs)."""
import os

import numpy as np
import scipy.sparse as sp

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden(object):
    def __init__(self, path):
        self._z = np.load(path)
        self.n = int(self._z["n"])
        self.seed = int(self._z["seed"])
        self.sim_num = int(self._z["sim_num"])

    def __getitem__(self, k):
        return self._z[k]

    def keys(self):
        return self._z.files

    def csr(self, prefix, shape=None):
        ip = self._z[prefix + "_indptr"]
        n = ip.size - 1
        return sp.csr_matrix((self._z[prefix + "_data"].copy(), self._z[prefix + "_indices"].copy(), ip.copy()),
                             shape=shape or (n, n))

    def csc(self, prefix):
        ip = self._z[prefix + "_indptr"]
        n = ip.size - 1
        return sp.csc_matrix((self._z[prefix + "_data"].copy(), self._z[prefix + "_indices"].copy(), ip.copy()),
                             shape=(n, n))

    def mats(self, tag):
        A, E, H = self.csr("A"), self.csr("E"), self.csr("H")
        eye = sp.eye(self.n).tocsr()
        return {"k1": [A], "k3": [A, E, H], "k2": [A, eye], "k4": [A, E, H, eye]}[tag]


def load_golden(name):
    return Golden(os.path.join(GOLDEN_DIR, name + ".npz"))


def rel_err(a, b):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def philox_normals(n, ncols, col_begin, seed, stream):
    """numpy restatement of the library's counter-based probe generator (csrc/chol.cu probe_normals_kernel):
    Philox4x32-10 keyed by the seed, counter (row, global column, stream), Box-Muller on two 53-bit uniforms."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
    mask = np.uint64(0xFFFFFFFF)
    i = np.repeat(np.arange(n, dtype=np.uint64), ncols)
    c = np.tile(np.arange(col_begin, col_begin + ncols, dtype=np.uint64), n)
    c0, c1 = i & mask, c & mask
    c2 = np.full(i.shape, np.uint64(stream) & mask, dtype=np.uint64)
    c3 = (np.uint64(stream >> 32) & mask) ^ (i >> np.uint64(32))
    k0, k1 = np.uint64(seed) & mask, np.uint64(seed >> 32) & mask
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0, k1 = (k0 + W0) & mask, (k1 + W1) & mask
    a = ((c0 << np.uint64(32)) | c1) >> np.uint64(11)
    b = ((c2 << np.uint64(32)) | c3) >> np.uint64(11)
    u1 = (a.astype(np.float64) + 0.5) * 2.0 ** -53
    u2 = (b.astype(np.float64) + 0.5) * 2.0 ** -53
    return (np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)).reshape(n, ncols), u1.reshape(n, ncols)
