"""Developer probe (not part of the product): times the phases of one REML evaluation and the HE kernels on a
simulated pedigree, with CUDA events.  Usage: python scripts/perf_probe.py N SF [REMOVE_FRAC] [K] [S]"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scilmm_b200 import pedigree as P  # noqa: E402
from scilmm_b200 import engine as E  # noqa: E402
import scilmm_b200.SparseCholesky  # noqa: E402,F401
S = sys.modules["scilmm_b200.SparseCholesky"]


def ev_time(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    n = int(sys.argv[1])
    sf = float(sys.argv[2])
    frac = float(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3] != "-" else None
    K = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    s = int(sys.argv[5]) if len(sys.argv) > 5 else 128
    t = time.time()
    ped = P.simulate_pedigree(n, sf, seed=0, remove_frac=frac)
    A, T, D, F = P.numerator(ped["rel"])
    H = P.household_matrix(ped["household"])
    keep, (A, H) = P.drop_unrelated(A, H)
    nn = A.shape[0]
    print("generated n=%d kept=%d nnz=%d in %.1fs (remove_frac=%g)" % (n, nn, A.nnz, time.time() - t, ped["remove_frac"]),
          flush=True)
    mats = [A]
    if K >= 3:
        mats.append(P.epistasis(A))
    mats.append(sp.eye(nn).tocsr())
    rng = np.random.default_rng(1)
    cov = np.hstack([rng.standard_normal((nn, 10)), np.ones((nn, 1))])
    y = rng.standard_normal(nn)
    chol = S.SparseCholesky(rng="device")
    t = time.time()
    ses = chol._session(mats, cov, y)
    st = ses.eng.stats()
    print("session setup %.1fs  order %.1fs symbolic %.1fs" % (time.time() - t, st["t_order"], st["t_symbolic"]))
    print({k: st[k] for k in ("nsuper", "nlevels", "nnzL", "lsize", "max_front_rows", "max_super_cols", "launches",
                              "device_bytes", "flops", "issued_flops")}, flush=True)
    sig = np.array([0.3, 0.2, 0.5]) if len(mats) == 3 else np.array([0.4, 0.6])

    def assemble():
        for k in range(ses.K):
            ses.eng.add_values(ses.map_ids[k], ses.matset.values_ptr(k), float(sig[k]), k == 0)

    t_as = ev_time(assemble)

    def fact():
        assemble()
        ses.eng.factorize()

    t_f = ev_time(fact) - t_as
    print("assemble %.2f ms   factorize %.1f ms   %.2f TFLOP/s (colcount^2)  %.2f TFLOP/s (issued)" %
          (t_as, t_f, st["flops"] / t_f / 1e9, st["issued_flops"] / t_f / 1e9), flush=True)
    print("logdet", ses.eng.logdet())
    for nr in (12, s):
        B = torch.randn(nn, nr, dtype=torch.float64, device="cuda")
        t_s = ev_time(lambda: ses.eng.solve_(B.clone()))
        bytes_ = 2 * (8 * st["lsize"]) + 4 * 8 * nn * nr
        print("solve nrhs=%d: %.2f ms  (%.0f GB/s on 2x panel bytes, %.2f TFLOP/s)" %
              (nr, t_s, bytes_ / t_s / 1e6, 4.0 * st["nnzL"] * nr / t_s / 1e9), flush=True)
    Z = torch.randn(nn, s, dtype=torch.float64, device="cuda")
    t_l = ev_time(lambda: ses.eng.lmul(Z))
    print("lmul nrhs=%d: %.2f ms" % (s, t_l))
    X = torch.randn(nn, s + 1, dtype=torch.float64, device="cuda")
    for k in range(ses.K):
        t_c = ev_time(lambda: ses.matset.coldot(k, X))
        nnz = ses.matset.nnz[k]
        print("coldot k=%d ncols=%d: %.2f ms  %.0f GB/s (12 nnz + 8 n ncols)" %
              (k, s + 1, t_c, (12.0 * nnz + 8.0 * nn * (s + 1)) / t_c / 1e6))
    t_e = ev_time(lambda: ses.evaluate(sig, True, s), reps=2)
    print("full REML evaluation (device rng): %.1f ms" % t_e, flush=True)
    # HE
    hm = [A, P.epistasis(A), H]
    ms = E.MatSet(hm)
    yd = E.to_device(y)
    t_h = ev_time(lambda: ms.he_moments_device(yd))
    byt = sum(12.0 * m.nnz + 4 * (nn + 1) for m in [A, H]) + 8.0 * A.nnz + 3 * 8 * nn
    print("HE moments K=3: %.3f ms   %.0f GB/s (fused-minimum bytes %.2f GB)" % (t_h, byt / t_h / 1e6, byt / 1e9))
    t0 = time.time()
    est = S.HE(hm, cov, y.copy())
    print("HE() public call %.3fs  est %s" % (time.time() - t0, est))


if __name__ == "__main__":
    main()
