"""One REML evaluation (after one warm-up) on the BASELINE 250K config + one HE moments pass at a reduced size:
the command the ncu launch list and the per-kernel `--set full` captures under profiles/ are taken from.
Usage: python scripts/profile_step.py [n] [sf] [remove_frac]"""
import os, sys
import numpy as np, scipy.sparse as sp, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from scilmm_b200 import engine as E, pedigree as P
import scilmm_b200.SparseCholesky
S = sys.modules["scilmm_b200.SparseCholesky"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
sf = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
A, H, cov, y, info = bench.make_inputs(n, sf, 10, with_household=True)
mats = [A, P.epistasis(A), sp.eye(A.shape[0]).tocsr()]
ys = y / y.std()
chol = S.SparseCholesky(rng="device"); ses = chol._session(mats, cov, ys); sig = np.array([0.3, 0.15, 0.55])
ses.evaluate(sig, True, 128)                 # warm-up: plans, workspaces, allocator
torch.cuda.synchronize()
E.launch_count(reset=True)
nll, grad = ses.evaluate(sig, True, 128)     # the profiled evaluation
torch.cuda.synchronize()
print("evaluate launches:", E.launch_count(reset=True), "nll", nll)
ms = E.MatSet([A, P.epistasis(A), H]); yd = E.to_device(ys)
ms.he_moments_device(yd); torch.cuda.synchronize()
E.launch_count(reset=True)
ms.he_moments_device(yd); torch.cuda.synchronize()
print("HE launches:", E.launch_count(reset=True))
