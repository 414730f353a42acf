"""TEST INFRASTRUCTURE ONLY — CPU stand-ins for the CHOLMOD Factor the reference obtains from
sksparse.cholmod.cholesky (reference scilmm/SparseCholesky.py:16-26).  scikit-sparse / CHOLMOD are
not installed in this image (requirements.txt:2 pins only `scikit-sparse>=0.4.3`, CHOLMOD unpinned),
so the Factor protocol the reference relies on is restated here on LAPACK / SuperLU:

    factor(b)        -> V^-1 b in the caller's ordering          (call sites :30,32,52,100,149,153)
    factor.logdet()  -> log det V                                 (:40)
    factor.L()       -> sparse lower L with L L' = V[P][:, P]     (:50)
    factor.P()       -> fill-reducing permutation, int32          (:93)

Both back-ends accept a fixed permutation so that the oracle can be forced onto the engine's P
(L is unique given P; the REML probe vectors depend on P through L).
"""
import numpy as np
import scipy.linalg as la
import scipy.sparse as sp
import scipy.sparse.linalg as spla


class NotPositiveDefinite(np.linalg.LinAlgError):
    pass


class DenseFactor(object):
    """Dense LAPACK Cholesky of V[P][:, P]; exact reference for n up to ~12K."""

    def __init__(self, V, perm=None):
        n = V.shape[0]
        self._perm = np.arange(n, dtype=np.int32) if perm is None else np.asarray(perm, dtype=np.int32)
        Vd = V.toarray() if sp.issparse(V) else np.asarray(V, dtype=float)
        Vp = Vd[self._perm][:, self._perm]
        try:
            self._Lp = la.cholesky(Vp, lower=True)
        except la.LinAlgError as e:
            raise NotPositiveDefinite(str(e))
        self._pinv = np.argsort(self._perm)

    def __call__(self, b):
        bp = np.asarray(b, dtype=float)[self._perm]
        x = la.cho_solve((self._Lp, True), bp)
        return x[self._pinv]

    def logdet(self):
        return 2.0 * np.sum(np.log(np.diag(self._Lp)))

    def L(self):
        return sp.csc_matrix(np.tril(self._Lp))

    def P(self):
        return self._perm.copy()


class SuperLUFactor(object):
    """Sparse LL' through SuperLU in symmetric mode without pivoting (= Cholesky up to scaling).

    perm=None lets SuperLU pick MMD(A'+A); otherwise V is pre-permuted and factored in NATURAL order.
    """

    def __init__(self, V, perm=None):
        V = sp.csc_matrix(V)
        n = V.shape[0]
        if perm is None:
            lu = spla.splu(V, permc_spec='MMD_AT_PLUS_A', diag_pivot_thresh=0.0,
                           options=dict(SymmetricMode=True))
            if not np.array_equal(lu.perm_r, lu.perm_c):
                raise RuntimeError("SuperLU pivoted; matrix is not safely SPD")
            self._perm = np.argsort(lu.perm_r).astype(np.int32)
            self._pre = None
        else:
            self._perm = np.asarray(perm, dtype=np.int32)
            Vp = V[self._perm][:, self._perm].tocsc()
            lu = spla.splu(Vp, permc_spec='NATURAL', diag_pivot_thresh=0.0,
                           options=dict(SymmetricMode=True))
            if not (np.array_equal(lu.perm_r, np.arange(n)) and np.array_equal(lu.perm_c, np.arange(n))):
                raise RuntimeError("SuperLU pivoted; matrix is not safely SPD")
            self._pre = np.argsort(self._perm)
        d = lu.U.diagonal()
        if np.any(d <= 0):
            raise NotPositiveDefinite("non-positive pivot")
        self._lu, self._d = lu, d

    def __call__(self, b):
        b = np.asarray(b, dtype=float)
        if self._pre is None:
            return self._lu.solve(b)
        return self._lu.solve(b[self._perm])[self._pre]

    def logdet(self):
        return float(np.sum(np.log(self._d)))

    def L(self):
        return (self._lu.L @ sp.diags(np.sqrt(self._d))).tocsc()

    def P(self):
        return self._perm.copy()


def dense_cholesky_func(perm=None):
    return lambda V: DenseFactor(V, perm)


def superlu_cholesky_func(perm=None):
    return lambda V: SuperLUFactor(V, perm)
