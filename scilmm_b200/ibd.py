"""IBD (numerator relationship) matrix on the GPU: drop-ins for the reference's scilmm/Matrices/Numerator.py.

    LD(rel, return_inbreeding_coefficient=False)   Numerator.py:5-34
    create_numerator(L, D)                         Numerator.py:37-38
    simple_numerator(rel)                          Numerator.py:41-43   -> (A, L, D)

`rel` is the boolean child -> parents matrix (parents precede children, at most two per individual).  The kernels
(csrc/ibd.cu) follow the reference's arithmetic operation by operation, so L, D, F and A are bit-identical to the
reference's output; the per-individual Python loop (12 s at n = 10,000 there) becomes a handful of launches.
There is no CPU fallback here (scilmm_b200.pedigree.numerator is the host generator used where no GPU exists).
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import engine as _eng
from ._lib import check, lib, np_ptr


class _Handle(object):
    def __init__(self, rel):
        _eng.require_cuda()
        rel = sp.csr_matrix(rel)
        if not rel.has_sorted_indices:
            rel = rel.sorted_indices()
        if rel.shape[0] != rel.shape[1]:
            raise ValueError("rel must be square")
        self.n = rel.shape[0]
        ip = np.ascontiguousarray(rel.indptr, dtype=np.int32)
        ix = np.ascontiguousarray(rel.indices, dtype=np.int32)
        h = C.c_void_p()
        check(lib().slmm_ibd_build(self.n, np_ptr(ip), np_ptr(ix), C.byref(h)))
        self._h = h
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int32(0)
        check(lib().slmm_ibd_sizes(h, C.byref(a), C.byref(b), C.byref(c)))
        self.nnzL, self.nnzA, self.nlevels = a.value, b.value, c.value

    def _csr(self, fn, nnz):
        ip = np.zeros(self.n + 1, dtype=np.int32)
        ix = np.zeros(nnz, dtype=np.int32)
        dt = np.zeros(nnz, dtype=np.float64)
        check(fn(self._h, np_ptr(ip), np_ptr(ix), np_ptr(dt)))
        m = sp.csr_matrix((dt, ix, ip), shape=(self.n, self.n))
        m.has_sorted_indices = True
        return m

    def L(self):
        return self._csr(lib().slmm_ibd_copy_L, self.nnzL)

    def A(self):
        return self._csr(lib().slmm_ibd_copy_A, self.nnzA)

    def DF(self):
        D, F = np.zeros(self.n), np.zeros(self.n)
        check(lib().slmm_ibd_copy_DF(self._h, np_ptr(D), np_ptr(F)))
        return D, F

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().slmm_ibd_destroy(self._h)
                self._h = None
        except Exception:
            pass


def LD(rel, return_inbreeding_coefficient=False):
    """reference Numerator.py:5-34.  Returns (L csr, D csr diagonal[, F])."""
    h = _Handle(rel)
    D, F = h.DF()
    D_full = sp.diags(D).tocsr()
    if return_inbreeding_coefficient:
        return h.L(), D_full, F
    return h.L(), D_full


def create_numerator(L, D):
    """reference Numerator.py:37-38 (host scipy product for callers that hold L and D already)."""
    return L.dot(D).dot(L.transpose(copy=True)).tocsr()


def simple_numerator(rel):
    """reference Numerator.py:41-43: (A, L, D) with A = L D L' built on the device in one call."""
    h = _Handle(rel)
    D, _ = h.DF()
    return h.A(), h.L(), sp.diags(D).tocsr()


def numerator(rel):
    """(A, T, D, F) like scilmm_b200.pedigree.numerator, computed on the GPU."""
    h = _Handle(rel)
    D, F = h.DF()
    return h.A(), h.L(), D, F
