"""Developer probe: per-launch profile of one factorization (and one 128-RHS solve) on a simulated pedigree."""
import os, sys, time
import numpy as np, scipy.sparse as sp, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from scilmm_b200 import pedigree as P
import scilmm_b200.SparseCholesky
S = sys.modules["scilmm_b200.SparseCholesky"]
n, sf = int(sys.argv[1]), float(sys.argv[2]); frac = float(sys.argv[3]) if len(sys.argv) > 3 else None
ped = P.simulate_pedigree(n, sf, seed=0, remove_frac=frac)
A, T, D, F = P.numerator(ped["rel"]); keep, (A,) = P.drop_unrelated(A); nn = A.shape[0]
mats = [A, P.epistasis(A), sp.eye(nn).tocsr()]
rng = np.random.default_rng(1); cov = np.hstack([rng.standard_normal((nn, 10)), np.ones((nn, 1))]); y = rng.standard_normal(nn)
chol = S.SparseCholesky(rng="device"); ses = chol._session(mats, cov, y); sig = np.array([0.3, 0.15, 0.55])
names = ["potrf", "gemm_big", "gemm_small", "extend_add", "rhs_pull", "extend_add_big", "init_w", "reduce", "-", "-", "skinny_f1", "skinny_f2", "gemm_tma"]
def report(tag):
    ms, fl, kind, grid = ses.eng.launch_profile()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez(os.path.join(ROOT, "gpurun_out", "launches_%s.npz" % tag.split()[0]), ms=ms, fl=fl, kind=kind, grid=grid)
    print("==", tag, "launches", ms.size, "total %.1f ms" % ms.sum())
    for k in range(13):
        m = kind == k
        if m.sum() == 0: continue
        print("  %-10s n=%5d  %.1f ms  %.2f TFLOP/s" % (names[k], m.sum(), ms[m].sum(), fl[m].sum() / max(ms[m].sum(), 1e-9) / 1e9))
    for k in (1, 2, 12):
        m = kind == k
        if m.sum() == 0: continue
        rate = fl[m] / np.maximum(ms[m], 1e-6) / 1e9
        for lo, hi in ((0, 2), (2, 5), (5, 10), (10, 15), (15, 20), (20, 25), (25, 40)):
            b = (rate >= lo) & (rate < hi)
            if b.sum():
                print("    %s rate %2d-%2d TF: n=%4d ms=%.1f flops=%.2e  median grid %d" % (names[k], lo, hi, b.sum(), ms[m][b].sum(), fl[m][b].sum(), np.median(grid[m][b])))
    top = np.argsort(-ms)[:12]
    print("  slowest launches:", [(names[kind[i]], round(float(ms[i]), 2), int(grid[i]), "%.1fTF" % (fl[i] / max(ms[i], 1e-6) / 1e9)) for i in top])
ses.factor_at(sig); ses.factor_at(sig)
for k in range(ses.K): ses.eng.add_values(ses.map_ids[k], ses.matset.values_ptr(k), float(sig[k]), k == 0)
ses.eng.set_profiling(True); ses.eng.factorize(); report("factorize"); ses.eng.set_profiling(False)
B = torch.randn(nn, 128, dtype=torch.float64, device="cuda"); ses.eng.solve_(B.clone())
ses.eng.set_profiling(True); ses.eng.solve_(B.clone()); report("solve 128 rhs"); ses.eng.set_profiling(False)
ses.eng.lmul(B); ses.eng.set_profiling(True); ses.eng.lmul(B); report("lmul 128"); ses.eng.set_profiling(False)
B12 = torch.randn(nn, 12, dtype=torch.float64, device="cuda"); ses.eng.solve_(B12.clone())
ses.eng.set_profiling(True); ses.eng.solve_(B12.clone()); report("solve12 12 rhs"); ses.eng.set_profiling(False)
