import sys; sys.path.insert(0,'.')
import numpy as np, scipy.sparse as sp, time
import scilmm_b200.SparseCholesky
S = sys.modules["scilmm_b200.SparseCholesky"]
n=int(sys.argv[1])
rng=np.random.default_rng(0); B=rng.standard_normal((n,n)); V=sp.csc_matrix(B@B.T+n*np.eye(n))
f=S.SparseCholesky(ordering_method="natural")(V)
f=S.SparseCholesky(ordering_method="natural")(V)
print(f.logdet(), np.linalg.slogdet(V.toarray())[1])
