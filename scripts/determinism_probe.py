"""Developer probe: is the factorization / solve bit-reproducible at the 250K config?"""
import os, sys
import numpy as np, scipy.sparse as sp, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from scilmm_b200 import engine as E, pedigree as P
import scilmm_b200.SparseCholesky
S = sys.modules["scilmm_b200.SparseCholesky"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
A, _, cov, y, info = bench.make_inputs(n, 1e-3, 10)
mats = [A, P.epistasis(A), sp.eye(A.shape[0]).tocsr()]
ys = y / y.std()
chol = S.SparseCholesky(rng="device"); ses = chol._session(mats, cov, ys); sig = np.array([0.3, 0.15, 0.55])
B = torch.from_numpy(np.random.default_rng(3).standard_normal((ses.n, 4))).cuda()
def one(tag, prof=False):
    ses.eng.set_profiling(prof)
    ses.factor_at(sig)
    ld = ses.eng.logdet()
    X = ses.eng.solve_(B.clone())
    ses.eng.set_profiling(False)
    torch.cuda.synchronize()
    print("%-12s logdet %.15e  x.sum %.15e  x[7].sum %.15e" % (tag, ld, float(X.sum()), float(X[7].sum())), flush=True)
    return ld, X
r = [one("streams %d" % i) for i in range(4)]
p = [one("serial %d" % i, True) for i in range(3)]
for i in range(1, 4): print("streams", i, "bitwise equal to run 0:", r[i][0] == r[0][0], bool(torch.equal(r[i][1], r[0][1])))
for i in range(3): print("serial", i, "vs streams 0:", p[i][0] == r[0][0], bool(torch.equal(p[i][1], r[0][1])), " max|dx| %.3e" % float((p[i][1] - r[0][1]).abs().max()))
# residual check: V x = b
V = (sig[0] * mats[0] + sig[1] * mats[1] + sig[2] * mats[2]).tocsr()
Xh = r[0][1].cpu().numpy()
res = V.dot(Xh) - B.cpu().numpy()
print("residual max |Vx-b| = %.3e  (|b| max %.2f)" % (np.abs(res).max(), float(B.abs().max())))
