"""Developer probe: dense SPD matrices (one supernode) of growing size - determinism and accuracy of factor/logdet."""
import sys, time; sys.path.insert(0, '.')
import numpy as np, scipy.sparse as sp, torch
import scilmm_b200.SparseCholesky
S = sys.modules["scilmm_b200.SparseCholesky"]
for n in [int(a) for a in sys.argv[1:]] or [1500, 3000, 6000, 10000]:
    rng = np.random.default_rng(0)
    B = rng.standard_normal((n, 64))
    Vd = 0.3 * (B @ B.T) / 64 + np.eye(n)
    V = sp.csc_matrix(Vd)
    chol = S.SparseCholesky(ordering_method="natural")
    lds = []
    for rep in range(3):
        f = chol(V)
        lds.append(f.logdet())
    ref = np.linalg.slogdet(Vd)[1]
    b = rng.standard_normal(n)
    x = f(b)
    print("n=%d logdets %s  ref %.12e  rel err %.2e  residual %.2e" % (n, ["%.12e" % v for v in lds], ref, abs(lds[-1] - ref) / abs(ref),
          np.abs(Vd @ x - b).max()), flush=True)
