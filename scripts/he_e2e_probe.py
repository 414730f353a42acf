"""Developer probe: where the end-to-end time of HE() at the 1M config goes (host prep, pinning, H2D, kernels)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from scilmm_b200 import engine as E, pedigree as P
import scilmm_b200.SparseCholesky
S = sys.modules["scilmm_b200.SparseCholesky"]
A, H, cov, y, info = bench.make_inputs(1000000, 1e-4, 2, seed=0, with_household=True)
mats = [A, P.epistasis(A), H]
print(info, "cpus", os.cpu_count(), flush=True)
torch.cuda.init(); torch.zeros(1, device="cuda")
def T(tag, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print("%-40s %8.1f ms" % (tag, (time.perf_counter() - t0) * 1e3), flush=True); return r
for rep in range(2):
    print("--- rep", rep)
    T("HE() public call", lambda: S.HE(list(mats), cov, y.copy()))
    cm = T("canonical_csr x3", lambda: [E.canonical_csr(m) for m in mats])
    T("array_equal indices A vs AoA", lambda: np.array_equal(cm[0].indices, cm[1].indices))
    t = T("from_numpy+pin data (849MB)", lambda: torch.from_numpy(cm[0].data).pin_memory())
    T("pinned -> cuda", lambda: t.to("cuda", non_blocking=True))
    T("pageable -> cuda (torch)", lambda: torch.from_numpy(cm[0].data).to("cuda"))
    T("MatSet()", lambda: E.MatSet(mats))
    T("cov regression + std (host)", lambda: (y - cov.dot(np.linalg.solve(cov.T.dot(cov), cov.T.dot(y)))).std())
