"""Developer probe: HE moments at the BASELINE 1M config (K=3) - a few calls for ncu launch lists."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from scilmm_b200 import engine as E, pedigree as P
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
sf = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-4
A, H, cov, y, info = bench.make_inputs(n, sf, 2, seed=0, with_household=True)
print(info, H.nnz, flush=True)
ms = E.MatSet([A, P.epistasis(A), H]); yd = E.to_device(y)
for _ in range(3): ms.he_moments_device(yd)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ms.he_moments_device(yd)
e1.record(); torch.cuda.synchronize()
print("he_moments %.3f ms" % (e0.elapsed_time(e1) / 10))
out = ms.he_moments_device(yd).cpu().numpy()
print("moments", np.array2string(out[:8], precision=17))
