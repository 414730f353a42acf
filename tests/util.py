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
