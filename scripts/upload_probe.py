"""Developer probe: host -> device rate of a 1.6 GB pageable numpy array through (a) the library's staged ring with
1/2/4/8 worker threads, (b) torch pin_memory() + copy, (c) plain torch copy from pageable memory."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from scilmm_b200 import engine as E
a = np.random.default_rng(0).standard_normal(200_000_000)      # 1.6 GB, touched
torch.cuda.synchronize()
def t(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); x = fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0); del x
    return best
for th in (1, 2, 4, 8):
    E.UPLOAD_THREADS = th
    dt = t(lambda: E.to_device(a))
    print("staged ring, %d threads: %.3f s  %.1f GB/s" % (th, dt, a.nbytes / dt / 1e9), flush=True)
dt = t(lambda: torch.from_numpy(a).pin_memory().to("cuda", non_blocking=True))
print("pin_memory + copy: %.3f s  %.1f GB/s" % (dt, a.nbytes / dt / 1e9))
dt = t(lambda: torch.from_numpy(a).to("cuda"))
print("pageable copy: %.3f s  %.1f GB/s" % (dt, a.nbytes / dt / 1e9))
p = torch.from_numpy(a).pin_memory()
dt = t(lambda: p.to("cuda", non_blocking=True))
print("already pinned: %.3f s  %.1f GB/s" % (dt, a.nbytes / dt / 1e9))
t0 = time.perf_counter(); b = a.copy(); print("host memcpy 1 thread: %.1f GB/s" % (a.nbytes / (time.perf_counter() - t0) / 1e9))
rt = torch.cuda.cudart()
t0 = time.perf_counter(); rc = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0); t1 = time.perf_counter()
print("cudaHostRegister rc", rc, "%.3f s (%.1f GB/s)" % (t1 - t0, a.nbytes / (t1 - t0) / 1e9))
dt = t(lambda: torch.from_numpy(a).to("cuda", non_blocking=True))
print("registered in place: %.3f s  %.1f GB/s" % (dt, a.nbytes / dt / 1e9))
t0 = time.perf_counter(); rt.cudaHostUnregister(a.ctypes.data); print("unregister %.3f s" % (time.perf_counter() - t0))
