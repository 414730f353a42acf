"""Host logic (no GPU): symbolic analysis checked by emulating the multifrontal plan in numpy, and the
C-ABI surface (every symbol declared in include/scilmm_b200.h is exported by the library)."""
import os
import re
import subprocess

import numpy as np
import pytest
import scipy.linalg as la
import scipy.sparse as sp

from scilmm_b200 import _lib
from scilmm_b200 import engine as E
from scilmm_b200.engine import SymbolicView

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def emulate_plan(V, sym):
    """Run the engine's plan (entry map -> parent-pull extend-add -> partial dense Cholesky per front, levels
    deepest first) with numpy; returns the dense permuted factor."""
    a = sym.arrays()
    n = sym.n
    Vc = sp.csr_matrix(V)
    Vc.sort_indices()
    Lx = np.zeros(sym.lsize)
    tgt = sym.entry_map(Vc)
    ok = tgt >= 0
    assert ok.sum() == (Vc.nnz + n) // 2            # one copy of every symmetric pair + the diagonal
    np.add.at(Lx, tgt[ok], Vc.data[ok])
    children = [[] for _ in range(sym.nsuper)]
    for s in range(sym.nsuper):
        if a['sn_parent'][s] >= 0:
            children[a['sn_parent'][s]].append(s)
    U = {}
    L = np.zeros((n, n))
    for lev in range(sym.nlevels - 1, -1, -1):
        for s in a['level_sn'][a['level_ptr'][lev]:a['level_ptr'][lev + 1]]:
            f = a['sn_first'][s]
            ns = a['sn_first'][s + 1] - f
            ms = a['sn_nrow'][s]
            rs = ms - ns
            rows = a['rows'][a['sn_rowptr'][s]:a['sn_rowptr'][s + 1]]
            assert np.array_equal(rows[:ns], np.arange(f, f + ns)) and np.all(np.diff(rows) > 0)
            ld = int(a['sn_ld'][s])          # panels are padded to an even number of rows
            P = Lx[a['sn_lptr'][s]:a['sn_lptr'][s + 1]].reshape(ns, ld).T[:ms].copy()
            F22 = np.zeros((rs, rs))
            for c in children[s]:
                nsc = a['sn_first'][c + 1] - a['sn_first'][c]
                relc = a['rel'][a['sn_rowptr'][c] + nsc:a['sn_rowptr'][c + 1]]
                Uc = U.pop(c)
                for tt in range(relc.size):
                    for u in range(tt, relc.size):
                        if relc[tt] < ns:
                            P[relc[u], relc[tt]] += Uc[u, tt]
                        else:
                            F22[relc[u] - ns, relc[tt] - ns] += Uc[u, tt]
            D = np.tril(P[:ns, :ns])
            D = D + np.tril(D, -1).T
            L11 = la.cholesky(D, lower=True)
            L21 = la.solve_triangular(L11, P[ns:, :].T, lower=True).T if rs else np.zeros((0, ns))
            U[s] = (np.tril(F22) + np.tril(F22, -1).T - L21 @ L21.T) if rs else np.zeros((0, 0))
            L[np.ix_(rows[:ns], np.arange(f, f + ns))] = L11
            if rs:
                L[np.ix_(rows[ns:], np.arange(f, f + ns))] = L21
    return L, a


@pytest.mark.parametrize("ordering", ["natural", "metis"])
@pytest.mark.parametrize("tag", ["k2", "k4"])
def test_plan_reproduces_dense_cholesky(golden_small, ordering, tag):
    V = golden_small.csc("V_" + tag)
    sym = SymbolicView(V, ordering=ordering)
    L, a = emulate_plan(V, sym)
    perm = a['perm']
    assert sorted(perm.tolist()) == list(range(sym.n))
    if ordering == "natural":
        assert np.array_equal(perm, np.arange(sym.n))
    Lref = la.cholesky(V.toarray()[perm][:, perm], lower=True)
    assert np.max(np.abs(L - Lref)) < 1e-12
    true_counts = (np.abs(Lref) > 0).sum(axis=0)
    assert np.array_equal(true_counts, a['colcount'])
    assert sym.nnzL == true_counts.sum()
    assert sym.flops == float((true_counts.astype(float) ** 2).sum())


def test_given_permutation_is_kept(golden_small):
    V = golden_small.csc("V_k2")
    rng = np.random.default_rng(5)
    perm = rng.permutation(golden_small.n).astype(np.int32)
    sym = SymbolicView(V, perm=perm)
    L, a = emulate_plan(V, sym)
    assert np.array_equal(a['perm'], perm)
    Lref = la.cholesky(V.toarray()[perm][:, perm], lower=True)
    assert np.max(np.abs(L - Lref)) < 1e-12


def test_edge_patterns():
    # diagonal matrix, one dense block, and a ragged mix with isolated vertices
    for M in (sp.eye(7).tocsr(), sp.csr_matrix(np.ones((9, 9)) + 9 * np.eye(9)),
              sp.block_diag([np.ones((5, 5)) + 5 * np.eye(5), np.eye(3), [[4.0, 1], [1, 4]]]).tocsr()):
        for ordering in ("natural", "metis"):
            sym = SymbolicView(M, ordering=ordering)
            L, a = emulate_plan(M, sym)
            p = a['perm']
            assert np.max(np.abs(L @ L.T - M.toarray()[p][:, p])) < 1e-12


def test_c1mini_sizes(golden_c1mini):
    V = golden_c1mini.csc("V_k4")
    sym = SymbolicView(V, ordering="metis")
    assert sym.nsuper < sym.n and sym.nlevels < 64
    assert sym.lsize >= sym.nnzL
    a = sym.arrays()
    # levels partition the supernodes and every child sits exactly one level below its parent
    depth = np.full(sym.nsuper, -1)
    for lev in range(sym.nlevels):
        depth[a['level_sn'][a['level_ptr'][lev]:a['level_ptr'][lev + 1]]] = lev
    assert np.all(depth >= 0)
    has_parent = a['sn_parent'] >= 0
    assert np.all(depth[has_parent] == depth[a['sn_parent'][has_parent]] + 1)


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "scilmm_b200.h")).read()
    declared = set(re.findall(r"\b(slmm_[a-z_A-Z0-9]+)\s*\(", header))
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(re.findall(r" T (slmm_\w+)", out))
    assert declared, "no declarations parsed"
    assert declared <= exported, "missing: %s" % sorted(declared - exported)
    assert set(_lib.SIGNATURES) == declared
    assert _lib.lib().slmm_version() == 100


def test_tuned_ordering_beats_metis_defaults(monkeypatch):
    """The engine's METIS options (UFACTOR 10, NSEPS 3) must not cost more factorization flops than METIS' own
    defaults on a pedigree pattern (guards the option indices against a different METIS build), and the experiment
    overrides SLMM_METIS_OPT / SLMM_RELAX must be honoured."""
    from scilmm_b200 import engine as E, pedigree as P
    ped = P.simulate_pedigree(50000, 1e-3, seed=0, remove_frac=0.103125)     # large enough for ND to matter
    A, T, D, F = P.numerator(ped["rel"])
    keep, (A,) = P.drop_unrelated(A)
    monkeypatch.delenv("SLMM_METIS_OPT", raising=False)
    monkeypatch.delenv("SLMM_RELAX", raising=False)
    tuned = E.SymbolicView(A, ordering="metis")
    monkeypatch.setenv("SLMM_METIS_OPT", "16:200,15:1")           # METIS' defaults for node ND
    stock = E.SymbolicView(A, ordering="metis")
    assert tuned.flops <= stock.flops          # 5.8e9 against 6.5e9 here; 4.4e12 against 5.1e12 at the 250K config
    assert tuned.flops != stock.flops                             # the override reached METIS
    monkeypatch.delenv("SLMM_METIS_OPT")
    monkeypatch.setenv("SLMM_RELAX", "0,0,0,0,0,0")               # no amalgamation: fundamental supernodes only
    fund = E.SymbolicView(A, ordering="metis")
    assert fund.nsuper > tuned.nsuper and fund.flops == tuned.flops and fund.lsize < tuned.lsize


@pytest.mark.parametrize("case,tag", [("golden_small", "k4"), ("golden_c1mini", "k4"), ("golden_c1", "k2")])
def test_independent_symbolic_agrees_with_product(case, tag, request):
    """oracle/symbolic_ref.py (textbook etree / row-subtree counts in oracle/cpu_kernels.c, no product code) against
    the product's host analysis on the same permutation: elimination tree, column counts, nnz(L), flops - and the
    LAPACK multifrontal factor built on the independent structure against dense Cholesky."""
    from oracle.cpu_factor import DenseFactor
    from oracle.supernodal_cpu import SupernodalCPUFactor
    from oracle.symbolic_ref import IndependentPlan
    g = request.getfixturevalue(case)
    V = g.csc("V_" + tag)
    sym = SymbolicView(V, ordering="metis")
    a = sym.arrays()
    ip = IndependentPlan(V, a['perm'])
    assert np.array_equal(ip.parent, a['parent'])
    assert np.array_equal(ip.colcount, a['colcount'])
    assert ip.sym.nnzL == sym.nnzL and ip.sym.flops == sym.flops
    assert ip.sym.nsuper >= sym.nsuper                   # no relaxed amalgamation on the oracle side
    f = SupernodalCPUFactor(V, plan=ip)
    if g.n <= 2000:
        d = DenseFactor(V, a['perm'])
        b = np.random.default_rng(0).standard_normal((g.n, 3))
        assert abs(f.logdet() - d.logdet()) < 1e-12 * abs(d.logdet())
        assert np.max(np.abs(f(b) - d(b))) < 1e-11 * np.max(np.abs(d(b)))
        assert abs(f.L() - d.L()).max() < 1e-12
    else:
        assert abs(f.logdet() - g["logdet_" + tag]) < 1e-11 * abs(g["logdet_" + tag])
        assert np.max(np.abs(f(g["y"] / g["y"].std()) - g["Viy_" + tag])) < 1e-10 * np.max(np.abs(g["Viy_" + tag]))


def test_entry_map_triangle_rules(golden_small):
    """tri = 0 (both triangles, mirrored copy skipped) and CHOLMOD's one-triangle rules tri = +1 / -1 place the same
    values at the same panel offsets for a symmetric matrix; the other triangle is ignored."""
    import ctypes as C
    from scilmm_b200._lib import check, lib, np_ptr
    V = golden_small.csc("V_k4")
    pat = E.canonical_csr(sp.csr_matrix((V.data, V.indices, V.indptr), shape=V.shape))
    sym = SymbolicView(pat, ordering="metis")
    panels = []
    for tri in (0, 1, -1):
        tgt = np.zeros(pat.nnz, dtype=np.int64)
        check(lib().slmm_symbolic_entry_map_tri(sym._h, np_ptr(pat.indptr), np_ptr(pat.indices), tri, np_ptr(tgt)))
        rows = np.repeat(np.arange(pat.shape[0]), np.diff(pat.indptr))
        if tri > 0:
            assert np.all((tgt >= 0) == (pat.indices >= rows))
        if tri < 0:
            assert np.all((tgt >= 0) == (pat.indices <= rows))
        Lx = np.zeros(sym.lsize)
        ok = tgt >= 0
        assert np.unique(tgt[ok]).size == ok.sum()
        Lx[tgt[ok]] = pat.data[ok]
        panels.append(Lx)
    assert np.array_equal(panels[0], panels[1]) and np.array_equal(panels[0], panels[2])


@pytest.mark.parametrize("case,tag", [("golden_small", "k4"), ("golden_c1mini", "k4")])
def test_cpu_supernodal_oracle_on_the_engine_plan(case, tag, request):
    """oracle/supernodal_cpu.py on the ENGINE's plan (padded panels, relaxed supernodes: the layout the panel-by-panel
    GPU comparison uses) against dense LAPACK: factor, logdet, solves, L*Z."""
    from oracle.cpu_factor import DenseFactor
    from oracle.supernodal_cpu import SupernodalCPUFactor, SupernodalPlan
    g = request.getfixturevalue(case)
    V = g.csc("V_" + tag)
    plan = SupernodalPlan(sp.csr_matrix((V.data, V.indices, V.indptr), shape=V.shape))
    assert np.any(plan.a["sn_ld"] != plan.a["sn_nrow"])        # some panel is padded
    f = SupernodalCPUFactor(V, plan=plan)
    d = DenseFactor(V, plan.a["perm"])
    b = np.random.default_rng(0).standard_normal((g.n, 4))
    assert abs(f.logdet() - d.logdet()) < 1e-12 * abs(d.logdet())
    assert np.max(np.abs(f(b) - d(b))) < 1e-11 * np.max(np.abs(d(b)))
    assert abs(f.L() - d.L()).max() < 1e-12
    assert np.max(np.abs(f.lmul_unperm(b) - d.L().dot(b)[np.argsort(d.P())])) < 1e-11


def test_threaded_pattern_permutation_matches_the_sequential_one():
    """csrc/symbolic.cpp permute_pattern has two formulations (sequential counting scatter; per-column gather + sort
    on host threads, chosen on machines with >= 6 cores): the whole analysis must come out identical, and the
    threaded one must refuse a structurally asymmetric pattern like the sequential one does."""
    import os, subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys, numpy as np, scipy.sparse as sp, hashlib
        sys.path.insert(0, %r)
        from scilmm_b200 import engine as E
        rng = np.random.default_rng(5)
        n = 6000
        B = sp.random(n, n, density=0.004, random_state=3, format='csr')
        A = (B + B.T + sp.eye(n)).tocsr(); A.data[:] = 1.0
        a = E.SymbolicView(A, ordering='metis').arrays()
        h = hashlib.sha256()
        for k in ('perm', 'parent', 'colcount', 'sn_first', 'sn_nrow', 'rows', 'rel', 'level_sn'):
            h.update(np.ascontiguousarray(a[k]).tobytes())
        print('HASH', h.hexdigest())
        U = sp.triu(A, 1).tolil(); U[3, 4000] = 1.0
        bad = (sp.tril(A) + U.tocsr()).tocsr()
        try:
            E.SymbolicView(bad, ordering='metis')
            print('ASYM accepted')
        except Exception as e:
            print('ASYM', 'symmetric' in str(e))
    """ % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    outs = []
    for t in ("1", "8"):
        env = dict(os.environ, SLMM_HOST_THREADS=t)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append([l for l in r.stdout.splitlines() if l.startswith(("HASH", "ASYM"))])
    assert outs[0] == outs[1] and outs[0][0].startswith("HASH") and outs[0][1] == "ASYM True", outs


def test_fast_nested_dissection_is_a_valid_cheaper_ordering():
    """ordering 'nesdis_fast' (one METIS separator per bisection instead of the best of three): a permutation, the
    same analysis invariants, and a flop count within 15 % of the default on a pedigree pattern."""
    import bench
    from scilmm_b200 import engine as E
    A, _, cov, y, info = bench.make_inputs(20000, 1e-3, 2)
    a = E.SymbolicView(A, ordering="nesdis")
    b = E.SymbolicView(A, ordering="nesdis_fast")
    pb = b.arrays()["perm"]
    assert np.array_equal(np.sort(pb), np.arange(A.shape[0]))
    assert b.nnzL >= A.shape[0] and 0.85 < b.flops / a.flops < 1.15
