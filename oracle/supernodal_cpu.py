"""TEST INFRASTRUCTURE ONLY - CPU supernodal (multifrontal) LL' with LAPACK/BLAS on all host cores.

This is the CPU baseline of record for problem sizes where dense LAPACK / SuperLU (oracle/cpu_factor.py) are
out of reach.  It restates what CHOLMOD's supernodal factorization does for the reference
(sksparse.cholmod.cholesky(mode='supernodal'), reference scilmm/SparseCholesky.py:22-26): dense
dpotrf / dtrsm / dsyrk on supernodal fronts (threaded BLAS) with scatter steps in between, and the
supernodal forward / backward solves behind factor(b) (reference :30,32,52,100).  CHOLMOD is not installed in
this image, so this port - not CHOLMOD - is what gets timed; every report says so.

It implements the Factor protocol (__call__, logdet, L, P).  The fill-reducing ordering and supernode
partition come from the engine's host symbolic analysis (METIS nested dissection, the reference's 'nesdis'),
so both arms factor the same permuted matrix; the numerics are independent (LAPACK here, DMMA tiles there).
Unlike the reference (:22-26 from :92) the analysis is NOT repeated per call when `symbolic` is passed in;
pass symbolic=None to pay for it every time as the reference does.
"""
import ctypes as C

import numpy as np
import scipy.linalg.blas as blas
import scipy.linalg.lapack as lapack
import scipy.sparse as sp

from oracle import build_oracle

_lib = None
_threads = {"cur": None, "max": None}
SMALL_WORK = 200000      # ms * ns below which BLAS runs single-threaded (thread fork/join costs more than the op)


def _set_blas_threads(big):
    """OpenBLAS wakes every thread even for tiny operands; tens of thousands of small fronts then cost ~1 ms
    each.  Small fronts run on one thread, large fronts on all host cores (what a tuned CHOLMOD build does
    through its BLAS)."""
    import os
    import threadpoolctl
    if _threads["max"] is None:
        _threads["max"] = os.cpu_count() or 1
    want = _threads["max"] if big else 1
    if _threads["cur"] != want:
        threadpoolctl.threadpool_limits(limits=want, user_api="blas")
        _threads["cur"] = want


def _clib():
    global _lib
    if _lib is None:
        h = C.CDLL(build_oracle.build())
        vp, i64 = C.c_void_p, C.c_int64
        h.oracle_extend_add.argtypes = [vp, i64, vp, vp, i64, i64, vp, i64]
        h.oracle_scatter_sub_rows.argtypes = [vp, vp, i64, vp, i64]
        h.oracle_gather_rows.argtypes = [vp, vp, i64, vp, i64]
        _lib = h
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class NotPositiveDefinite(np.linalg.LinAlgError):
    pass


def _panel(Lx, a, s):
    """(ms x ns) view of panel s: column-major with leading dimension a['sn_ld'][s] (the engine pads panels to an even
    number of rows; an independent plan has ld = ms)."""
    ns = a['sn_first'][s + 1] - a['sn_first'][s]
    ms, ld, p0 = a['sn_nrow'][s], int(a['sn_ld'][s]), a['sn_lptr'][s]
    return Lx[p0:p0 + ld * ns].reshape((ld, ns), order='F')[:ms]


class SupernodalPlan(object):
    """Symbolic structures shared by factorizations with one pattern (host arrays from the engine's analysis)."""

    def __init__(self, pattern, ordering="metis", perm=None):
        from scilmm_b200.engine import SymbolicView, canonical_csr
        self.pattern = canonical_csr(pattern)
        self.sym = SymbolicView(self.pattern, ordering=ordering, perm=perm)
        self.a = self.sym.arrays()
        self.n = self.sym.n
        a = self.a
        self.children = [[] for _ in range(self.sym.nsuper)]
        for s in range(self.sym.nsuper):
            if a['sn_parent'][s] >= 0:
                self.children[a['sn_parent'][s]].append(s)
        self._maps = {}

    def entry_map(self, m):
        key = (m.nnz, m.indptr.ctypes.data, m.indices.ctypes.data)
        hit = self._maps.get(key)
        if hit is None:
            tgt = self.sym.entry_map(m)
            ok = np.flatnonzero(tgt >= 0)
            hit = (ok, tgt[ok], m)
            self._maps = {key: hit}
        return hit


class SupernodalCPUFactor(object):
    """plan: a SupernodalPlan (symbolic borrowed from the engine: same ordering and supernodes, panel-by-panel
    comparable) or an oracle.symbolic_ref.IndependentPlan (symbolic computed without the product library)."""

    def __init__(self, V, plan=None, ordering="metis", perm=None):
        from oracle.symbolic_ref import _csr32 as canonical_csr
        Vc = sp.csr_matrix((V.data, V.indices, V.indptr), shape=V.shape) if sp.isspmatrix_csc(V) else V
        Vc = canonical_csr(Vc)
        if plan is None:
            plan = SupernodalPlan(Vc, ordering=ordering, perm=perm)
        self.plan = plan
        a, sym = plan.a, plan.sym
        lib = _clib()
        ok, tgt, _ = plan.entry_map(Vc)
        Lx = np.zeros(sym.lsize)
        Lx[tgt] = Vc.data[ok]
        self.Lx = Lx
        first, nrow, lptr, rowptr = a['sn_first'], a['sn_nrow'], a['sn_lptr'], a['sn_rowptr']
        rel_all = a['rel']
        U = {}
        for lev in range(sym.nlevels - 1, -1, -1):
            for s in a['level_sn'][a['level_ptr'][lev]:a['level_ptr'][lev + 1]]:
                f = first[s]
                ns = first[s + 1] - f
                ms = nrow[s]
                rs = ms - ns
                panel = _panel(Lx, a, s)
                _set_blas_threads(ms * ns > SMALL_WORK)
                Up = np.zeros((rs, rs), order='F') if rs else None
                for c in plan.children[s]:
                    Uc = U.pop(c)
                    if Uc is None:
                        continue
                    nsc = first[c + 1] - first[c]
                    rel = rel_all[rowptr[c] + nsc:rowptr[c + 1]]
                    lib.oracle_extend_add(_p(Uc), Uc.shape[0], _p(rel), _p(panel), int(a['sn_ld'][s]), ns,
                                          _p(Up) if rs else None, rs)
                if ns == 1:
                    d = panel[0, 0]
                    if not d > 0:
                        raise NotPositiveDefinite("non-positive pivot at column %d" % f)
                    r = np.sqrt(d)
                    panel[0, 0] = r
                    if rs:
                        panel[1:, 0] /= r
                else:
                    c11, info = lapack.dpotrf(panel[:ns, :], lower=1, clean=1, overwrite_a=0)
                    if info != 0:
                        raise NotPositiveDefinite("non-positive pivot at column %d" % (f + info - 1))
                    panel[:ns, :] = c11
                    if rs:
                        panel[ns:, :] = blas.dtrsm(1.0, c11, panel[ns:, :], side=1, lower=1, trans_a=1)
                if rs:
                    L21 = panel[ns:, :]
                    U[s] = blas.dsyrk(-1.0, L21, beta=1.0, c=Up, lower=1, overwrite_c=1)
                else:
                    U[s] = None
        _set_blas_threads(True)

    def P(self):
        return self.plan.a['perm'].copy()

    def logdet(self):
        a = self.plan.a
        first, nrow, lptr = a['sn_first'], a['sn_nrow'], a['sn_lptr']
        tot = 0.0
        for s in range(self.plan.sym.nsuper):
            ns = first[s + 1] - first[s]
            ms = nrow[s]
            tot += np.sum(np.log(np.diagonal(_panel(self.Lx, a, s))[:ns]))
        return 2.0 * tot

    def L(self):
        a = self.plan.a
        n = self.plan.n
        first, nrow, lptr, rowptr = a['sn_first'], a['sn_nrow'], a['sn_lptr'], a['sn_rowptr']
        colptr = np.zeros(n + 1, dtype=np.int64)
        ri, vv = [], []
        for s in range(self.plan.sym.nsuper):
            f = first[s]
            ns = first[s + 1] - f
            ms = nrow[s]
            rows = a['rows'][rowptr[s]:rowptr[s + 1]]
            panel = _panel(self.Lx, a, s)
            for c in range(ns):
                ri.append(rows[c:])
                vv.append(panel[c:, c])
                colptr[f + c + 1] = ms - c
        colptr = np.cumsum(colptr)
        return sp.csc_matrix((np.concatenate(vv), np.concatenate(ri), colptr), shape=(n, n))

    def _sweeps(self, Xp, forward=True, backward=True):
        """Supernodal forward / backward substitution on a C-ordered (n x k) block.  BLAS sees the block as its
        transpose (k x n, Fortran order), so the rows of one supernode are a contiguous column block and the
        triangular solves run in place without copies."""
        a = self.plan.a
        lib = _clib()
        first, nrow, lptr, rowptr = a['sn_first'], a['sn_nrow'], a['sn_lptr'], a['sn_rowptr']
        nsuper = self.plan.sym.nsuper
        k = Xp.shape[1]
        XT = Xp.T                                          # (k x n) F-contiguous view
        if forward:
            for s in range(nsuper):                       # ascending ids = children before parents
                f = first[s]
                ns = first[s + 1] - f
                ms = nrow[s]
                panel = _panel(self.Lx, a, s)
                _set_blas_threads(ms * ns > SMALL_WORK)
                xt = XT[:, f:f + ns]
                if ns == 1:
                    xt /= panel[0, 0]
                else:                                      # x' := x' L11^-T
                    blas.dtrsm(1.0, panel[:ns, :], xt, side=1, lower=1, trans_a=1, overwrite_b=1)
                if ms > ns:
                    u = np.ascontiguousarray((xt @ panel[ns:, :].T).T)      # (rs x k) C-order
                    rows = a['rows'][rowptr[s] + ns:rowptr[s + 1]]
                    lib.oracle_scatter_sub_rows(_p(Xp), _p(rows), rows.size, _p(u), k)
        if backward:
            for s in range(nsuper - 1, -1, -1):
                f = first[s]
                ns = first[s + 1] - f
                ms = nrow[s]
                panel = _panel(self.Lx, a, s)
                _set_blas_threads(ms * ns > SMALL_WORK)
                xt = XT[:, f:f + ns]
                if ms > ns:
                    rows = a['rows'][rowptr[s] + ns:rowptr[s + 1]]
                    gb = np.empty((rows.size, k))
                    lib.oracle_gather_rows(_p(Xp), _p(rows), rows.size, _p(gb), k)
                    xt -= gb.T @ panel[ns:, :]
                if ns == 1:
                    xt /= panel[0, 0]
                else:                                      # x' := x' L11^-1
                    blas.dtrsm(1.0, panel[:ns, :], xt, side=1, lower=1, trans_a=0, overwrite_b=1)
        return Xp

    def __call__(self, b):
        b = np.asarray(b, dtype=np.float64)
        one = b.ndim == 1
        perm = self.plan.a['perm']
        Xp = np.ascontiguousarray((b[:, None] if one else b)[perm])
        self._sweeps(Xp)
        _set_blas_threads(True)
        out = np.empty_like(Xp)
        out[perm] = Xp
        return out[:, 0] if one else out

    def lmul_unperm(self, Z):
        """(L Z)[argsort P] without forming the simplicial L (what factor.L().dot(Z)[p_inv] computes)."""
        a = self.plan.a
        first, nrow, lptr, rowptr = a['sn_first'], a['sn_nrow'], a['sn_lptr'], a['sn_rowptr']
        Z = np.ascontiguousarray(Z, dtype=np.float64)
        out = np.zeros_like(Z)
        for s in range(self.plan.sym.nsuper):
            f = first[s]
            ns = first[s + 1] - f
            ms = nrow[s]
            rows = a['rows'][rowptr[s]:rowptr[s + 1]]
            panel = _panel(self.Lx, a, s)
            _set_blas_threads(ms * ns > SMALL_WORK)
            out[rows] += panel @ Z[f:f + ns]
        _set_blas_threads(True)
        res = np.empty_like(out)
        res[a['perm']] = out
        return res


def supernodal_cholesky_func(plan_cache=None, ordering="metis", reanalyze=False):
    """cholesky_func for the oracle's REML functions.  reanalyze=True repeats the symbolic analysis on every
    call like the reference does (SparseCholesky.py:22-26 from :92)."""
    cache = plan_cache if plan_cache is not None else {}

    def func(V):
        if reanalyze:
            return SupernodalCPUFactor(V, ordering=ordering)
        key = (V.shape[0], V.nnz)
        if key not in cache:
            Vc = sp.csr_matrix((V.data, V.indices, V.indptr), shape=V.shape) if sp.isspmatrix_csc(V) else V
            cache[key] = SupernodalPlan(Vc, ordering=ordering)
        return SupernodalCPUFactor(V, plan=cache[key])
    return func
