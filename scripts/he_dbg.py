import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from tests.util import load_golden
import scilmm_b200.SparseCholesky
S = sys.modules["scilmm_b200.SparseCholesky"]
g = load_golden("case_c1mini")
mats = g.mats("k3")
print([np.diff(m.indptr).max() for m in mats])
y = np.random.default_rng(1).standard_normal(g.n)
try:
    print(S.he_moments(mats, y)[:2])
except Exception as e:
    print("ERR", repr(e))
