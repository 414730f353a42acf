"""Developer probe: compare the GPU factor panel by panel with the CPU supernodal oracle (same analysis)."""
import os, sys
import numpy as np, scipy.sparse as sp, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from scilmm_b200 import engine as E, pedigree as P
import scilmm_b200.SparseCholesky
S = sys.modules["scilmm_b200.SparseCholesky"]
from oracle.supernodal_cpu import SupernodalCPUFactor, SupernodalPlan
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
A, _, cov, y, info = bench.make_inputs(n, 1e-3, 10)
mats = [A, P.epistasis(A), sp.eye(A.shape[0]).tocsr()]
sig = np.array([0.3, 0.15, 0.55])
chol = S.SparseCholesky(rng="device"); ses = chol._session(mats, cov, y / y.std())
V = (sig[0] * mats[0] + sig[1] * mats[1] + sig[2] * mats[2]).tocsr()
plan = SupernodalPlan(ses.union, perm=ses.eng.perm())
ref = SupernodalCPUFactor(V, plan=plan)
a = plan.a
first, nrow, lptr, par = a['sn_first'], a['sn_nrow'], a['sn_lptr'], a['sn_parent']
nsuper = len(nrow)
nch = np.bincount(par[par >= 0], minlength=nsuper)
lev = np.zeros(nsuper, int)
for d in range(len(a['level_ptr']) - 1):
    lev[a['level_sn'][a['level_ptr'][d]:a['level_ptr'][d + 1]]] = d
for rep in range(2):
    ses.eng.set_profiling(rep == 1)
    ses.factor_at(sig)
    ses.eng.set_profiling(False)
    Lx = ses.eng.panels()
    bad = []
    for s in range(nsuper):
        ns, ms = first[s + 1] - first[s], nrow[s]
        ld = int(a['sn_ld'][s])
        g = Lx[lptr[s]:lptr[s] + ld * ns].reshape((ld, ns), order='F')[:ms]
        r = ref.Lx[lptr[s]:lptr[s] + ld * ns].reshape((ld, ns), order='F')[:ms]
        m = np.tril(np.ones((min(ms, ns), ns), bool))
        diff = np.abs(g - r)
        diff[:ns][~m[:ns]] = 0
        e = diff.max()
        if e > 1e-9:
            i, j = np.unravel_index(np.argmax(diff), diff.shape)
            nbad = int((diff > 1e-9).sum())
            badcols = np.flatnonzero((diff > 1e-9).any(axis=0)); badrows = np.flatnonzero((diff > 1e-9).any(axis=1))
            bad.append((lev[s], s, ns, ms, int(nch[s]), e, nbad, (badcols.min(), badcols.max()), (badrows.min(), badrows.max())))
    bad.sort(key=lambda t: -t[0])
    print("rep", rep, "serial" if rep else "streams", "wrong supernodes:", len(bad), "of", nsuper)
    for t in bad[:12]:
        print("   level %d sn %d ns=%d ms=%d children=%d maxerr %.2e nbad=%d badcols %s badrows %s" % t)
