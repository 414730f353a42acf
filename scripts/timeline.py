"""Developer probe: when does each panel (chain stream) / bulk update (bulk stream) of the factorization finish?
Events of the two-stream schedule with timestamps (slmm_chol_set_timeline).  Usage: python scripts/timeline.py [n]"""
import os, sys
import numpy as np, scipy.sparse as sp, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from scilmm_b200 import pedigree as P
import scilmm_b200.SparseCholesky
S = sys.modules["scilmm_b200.SparseCholesky"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
A, _, cov, y, info = bench.make_inputs(n, 1e-3, 10)
mats = [A, P.epistasis(A), sp.eye(A.shape[0]).tocsr()]
chol = S.SparseCholesky(rng="device"); ses = chol._session(mats, cov, y / y.std()); sig = np.array([0.3, 0.15, 0.55])
ses.factor_at(sig); ses.factor_at(sig)
for k in range(ses.K): ses.eng.add_values(ses.map_ids[k], ses.matset.values_ptr(k), float(sig[k]), k == 0)
ms, st = ses.eng.timeline()
print("events", ms.size, "join at %.2f ms" % ms[1])
prev = {0: 0.0, 1: 0.0}
for q in range(2, ms.size):
    print("ev %4d stream %d  t=%8.2f ms  (+%.2f since previous on this stream)" % (q - 2, st[q], ms[q], ms[q] - prev[int(st[q])]))
    prev[int(st[q])] = ms[q]
end, stt, kind, grid, fl = ses.eng.launch_timeline()
names = ["potrf", "gemm_big", "gemm_small", "extend_add", "rhs_pull", "extend_add_big", "init_w", "reduce"]
np.savez(os.path.join(ROOT, "gpurun_out", "launch_timeline.npz"), end=end, stream=stt, kind=kind, grid=grid, flops=fl)
# per-stream durations: a launch starts no earlier than the previous launch on its stream ended
print("\nper 5 ms window: flops finished (TF) | chain-stream busy estimate")
T = float(end.max())
for w0 in np.arange(0, T, 5.0):
    m = (end >= w0) & (end < w0 + 5.0)
    print("  %6.1f-%6.1f ms: %6.2f TFLOP/s   launches: chain %3d bulk %3d" % (w0, w0 + 5, fl[m].sum() / 5e9, (m & (stt == 0)).sum(), (m & (stt == 1)).sum()))
prev = {0: 0.0, 1: 0.0}
gaps = []
for q in range(end.size):
    s_ = int(stt[q]); d = end[q] - prev[s_]; prev[s_] = end[q]
    gaps.append(d)
gaps = np.array(gaps)
for s_ in (0, 1):
    for k in range(8):
        m = (stt == s_) & (kind == k)
        if m.sum():
            print("stream %d %-14s n=%4d  sum(end-prev_end)=%.1f ms  flops=%.2e" % (s_, names[k], m.sum(), gaps[m].sum(), fl[m].sum()))
print("longest chain-stream steps:", sorted([(round(float(gaps[q]), 2), names[kind[q]], int(grid[q]), round(float(end[q]), 1)) for q in np.where(stt == 0)[0]], reverse=True)[:25])
