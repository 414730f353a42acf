"""Developer probe: where the e2e time of HE() goes at the 1M config - uploads of the individual CSR arrays, MatSet
construction (uploads + device-side validation / pattern comparison), first moments pass (symmetry, maps) and a warm one."""
import os, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from scilmm_b200 import engine as E, pedigree as P
A, H, cov, y, info = bench.make_inputs(1000000, 1e-4, 2, with_household=True)
AoA = P.epistasis(A)
E.UPLOAD_THREADS = 8
yd = E.to_device(y)
def tm(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%-46s %.4f s" % (label, dt), flush=True)
    return r, dt
for name, arr in (("A.indices", A.indices), ("A.data", A.data), ("AoA.indices", AoA.indices), ("AoA.data", AoA.data)):
    for rep in range(2):
        x, dt = tm("to_device %s (%d MB) rep %d" % (name, arr.nbytes >> 20, rep), lambda: E.to_device(arr))
        print("      %.1f GB/s" % (arr.nbytes / dt / 1e9)); del x
for rep in range(3):
    ms, dt = tm("MatSet([A, AoA, H]) rep %d" % rep, lambda: E.MatSet([A, AoA, H]))
    print("      %.1f GB/s on %d MB" % (ms.h2d_bytes / dt / 1e9, ms.h2d_bytes >> 20))
    tm("   first he_moments_device", lambda: ms.he_moments_device(yd).cpu())
    tm("   second he_moments_device", lambda: ms.he_moments_device(yd).cpu())
    del ms
torch.empty(900_000_000, dtype=torch.uint8, device="cuda"); torch.cuda.synchronize()
torch.cuda.empty_cache()
tm("torch.empty 850 MB (fresh cudaMalloc)", lambda: torch.empty(850_000_000, dtype=torch.uint8, device="cuda"))
