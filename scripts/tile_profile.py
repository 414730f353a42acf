"""Developer probe: per-CTA cycle counts of the tiled quadratic-form pass at the bench size, and a least-squares fit
of cycles ~ a*entries + b*rows + c*tiles (what the partition's cost model in quadform_tiled.cu is set from)."""
import os, sys, numpy as np, scipy.sparse as sp, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from scilmm_b200 import engine as E, pedigree as P
import scilmm_b200.SparseCholesky
S = sys.modules["scilmm_b200.SparseCholesky"]
A, _, cov, y, info = bench.make_inputs(250000, 1e-3, 10)
n = A.shape[0]
mats = [A, P.epistasis(A), sp.eye(n).tocsr()]
chol = S.SparseCholesky(rng="device")
ses = chol._session(mats, cov, y / y.std())
ms = ses.matset
for ncols, nb in ((140, 12), (128, 0), (28, 12)):
    X = torch.randn(n, ncols, dtype=torch.float64, device="cuda")
    for _ in range(3): ms.quadform_tiled([0, 1], X, nb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ms.quadform_tiled([0, 1], X, nb); e1.record(); torch.cuda.synchronize()
    prof = ms.tile_cta_profile(0).astype(np.float64)
    cyc = prof[:, 3]
    Xf = prof[:, [2, 1, 0]]
    coef, *_ = np.linalg.lstsq(Xf, cyc, rcond=None)
    pred = Xf @ coef
    print("ncols %d nb %d: %.2f ms; CTA cycles mean %.0f max %.0f min %.0f (max/mean %.2f); fit cycles = %.1f*entries + %.1f*rows + %.1f*tiles (rms rel err %.3f)"
          % (ncols, nb, e0.elapsed_time(e1), cyc.mean(), cyc.max(), cyc.min(), cyc.max() / cyc.mean(), coef[0], coef[1], coef[2],
             np.sqrt(np.mean(((pred - cyc) / cyc) ** 2))))
    print("   in units of one entry: per row %.1f, per tile %.1f" % (coef[1] / coef[0], coef[2] / coef[0]))
    i = int(np.argmax(cyc)); print("   slowest CTA: tiles %d rows %d entries %d cycles %d" % tuple(prof[i]))
np.save(os.path.join(ROOT, "gpurun_out", "tile_cta_profile.npy"), prof)
