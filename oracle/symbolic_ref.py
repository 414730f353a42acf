"""TEST INFRASTRUCTURE ONLY - an independent symbolic analysis for the parity tests.

The CPU supernodal factor (oracle/supernodal_cpu.py) normally borrows ordering, supernodes and scatter maps from
the product library (csrc/symbolic.cpp), so a symbolic bug would be shared by both arms.  This module restates the
analyze step CHOLMOD performs for the reference (sksparse.cholmod.cholesky, reference scilmm/SparseCholesky.py:22-26)
WITHOUT touching libscilmm_b200.so: textbook algorithms in oracle/cpu_kernels.c (Liu's elimination tree, column
counts and row structures by row-subtree traversal) + numpy glue.  Only the permutation is an input (any permutation
is valid; L is unique given P), typically the engine's factor.P().

IndependentPlan has the interface SupernodalCPUFactor expects from SupernodalPlan, so the same LAPACK multifrontal
numerics run on a completely separate symbolic structure: nnz(L), column counts, logdet and solves at the BASELINE
sizes are then cross-checked product-free.
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from oracle import build_oracle

_lib = None


def _clib():
    global _lib
    if _lib is None:
        h = C.CDLL(build_oracle.build())
        vp, i64 = C.c_void_p, C.c_int64
        h.oracle_etree.argtypes = [i64, vp, vp, vp]
        h.oracle_colcounts.argtypes = [i64, vp, vp, vp, vp]
        h.oracle_supernode_rows.argtypes = [i64, vp, vp, vp, vp, vp, i64, vp]
        h.oracle_entry_map.argtypes = [i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        h.oracle_entry_map.restype = C.c_int32
        _lib = h
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _csr32(m):
    m = sp.csr_matrix(m)
    if not m.has_sorted_indices:
        m = m.sorted_indices()
    return sp.csr_matrix((m.data.astype(np.float64), m.indices.astype(np.int32), m.indptr.astype(np.int32)),
                         shape=m.shape)


class _Sym(object):
    pass


class IndependentPlan(object):
    """etree / column counts / maximal supernodes / row structures / levels for pattern[perm][:, perm]."""

    def __init__(self, pattern, perm):
        lib = _clib()
        self.pattern = _csr32(pattern)
        n = self.n = self.pattern.shape[0]
        perm = np.ascontiguousarray(perm, dtype=np.int32)
        assert np.array_equal(np.sort(perm), np.arange(n)), "perm must be a permutation"
        iperm = np.empty(n, dtype=np.int32)
        iperm[perm] = np.arange(n, dtype=np.int32)
        ones = sp.csr_matrix((np.ones(self.pattern.nnz), self.pattern.indices, self.pattern.indptr), shape=(n, n))
        B = _csr32(ones[perm][:, perm])
        bp, bi = np.ascontiguousarray(B.indptr), np.ascontiguousarray(B.indices)
        parent = np.empty(n, dtype=np.int32)
        lib.oracle_etree(n, _p(bp), _p(bi), _p(parent))
        cc = np.empty(n, dtype=np.int32)
        lib.oracle_colcounts(n, _p(bp), _p(bi), _p(parent), _p(cc))
        self.parent, self.colcount = parent, cc
        # maximal supernodes: column j continues the supernode of j-1 when parent[j-1] == j and the structures nest
        join = np.zeros(n, dtype=bool)
        if n > 1:
            join[1:] = (parent[:-1] == np.arange(1, n)) & (cc[:-1] == cc[1:] + 1)
        firsts = np.flatnonzero(~join).astype(np.int32)
        nsuper = firsts.size
        sn_first = np.concatenate((firsts, [n])).astype(np.int32)
        col2sn = (np.cumsum(~join) - 1).astype(np.int32)
        sn_nrow = cc[firsts].astype(np.int32)
        sn_rowptr = np.concatenate(([0], np.cumsum(sn_nrow, dtype=np.int64))).astype(np.int64)
        ncols = np.diff(sn_first).astype(np.int64)
        sn_lptr = np.concatenate(([0], np.cumsum(sn_nrow.astype(np.int64) * ncols))).astype(np.int64)
        first_of = np.full(n, -1, dtype=np.int32)
        first_of[firsts] = np.arange(nsuper, dtype=np.int32)
        rows = np.empty(int(sn_rowptr[-1]), dtype=np.int32)
        lib.oracle_supernode_rows(n, _p(bp), _p(bi), _p(parent), _p(first_of), _p(sn_rowptr), nsuper, _p(rows))
        last = sn_first[1:] - 1
        sn_parent = np.where(parent[last] >= 0, col2sn[np.maximum(parent[last], 0)], -1).astype(np.int32)
        # relative indices of the below-diagonal rows inside the parent's row list
        rel = np.full(rows.size, -1, dtype=np.int32)
        for s in range(nsuper):
            p = sn_parent[s]
            if p < 0:
                continue
            a0, a1 = sn_rowptr[s] + ncols[s], sn_rowptr[s + 1]
            prow = rows[sn_rowptr[p]:sn_rowptr[p + 1]]
            pos = np.searchsorted(prow, rows[a0:a1])
            assert np.array_equal(prow[pos], rows[a0:a1]), "row structure is not nested in the parent"
            rel[a0:a1] = pos
        depth = np.zeros(nsuper, dtype=np.int32)
        for s in range(nsuper - 1, -1, -1):
            depth[s] = 0 if sn_parent[s] < 0 else depth[sn_parent[s]] + 1
        nlevels = int(depth.max()) + 1 if nsuper else 0
        order = np.argsort(depth, kind="stable").astype(np.int32)
        level_ptr = np.concatenate(([0], np.cumsum(np.bincount(depth, minlength=nlevels)))).astype(np.int32)
        self.a = dict(perm=perm, parent=parent, colcount=cc, sn_first=sn_first, sn_nrow=sn_nrow, sn_parent=sn_parent,
                      sn_rowptr=sn_rowptr, sn_lptr=sn_lptr, rows=rows, rel=rel, level_ptr=level_ptr, level_sn=order,
                      sn_ld=sn_nrow.astype(np.int64))
        self._iperm, self._col2sn = iperm, col2sn
        sym = _Sym()
        sym.n, sym.nsuper, sym.nlevels = n, nsuper, nlevels
        sym.lsize = int(sn_lptr[-1])
        sym.nnzL = int(cc.astype(np.int64).sum())
        sym.flops = float((cc.astype(np.float64) ** 2).sum())
        self.sym = sym
        self.children = [[] for _ in range(nsuper)]
        for s in range(nsuper):
            if sn_parent[s] >= 0:
                self.children[sn_parent[s]].append(s)
        self._maps = {}

    def entry_map(self, m):
        key = (m.nnz, m.indptr.ctypes.data, m.indices.ctypes.data)
        hit = self._maps.get(key)
        if hit is None:
            a = self.a
            tgt = np.empty(m.nnz, dtype=np.int64)
            rc = _clib().oracle_entry_map(self.n, _p(np.ascontiguousarray(m.indptr, dtype=np.int32)),
                                          _p(np.ascontiguousarray(m.indices, dtype=np.int32)), _p(self._iperm),
                                          _p(self._col2sn), _p(a['sn_first']), _p(a['sn_nrow']), _p(a['sn_rowptr']),
                                          _p(a['sn_lptr']), _p(a['rows']), _p(tgt))
            if rc != 0:
                raise ValueError("matrix entry outside the analysed pattern")
            ok = np.flatnonzero(tgt >= 0)
            hit = (ok, tgt[ok], m)
            self._maps = {key: hit}
        return hit
