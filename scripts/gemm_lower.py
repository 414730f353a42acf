import sys; sys.path.insert(0, '.')
from scilmm_b200 import engine as E
M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
err, ms, tf = E.gemm_selftest(M, N, K, lower=True, reps=2)
print(M, N, K, 'lower err %.2e  %.3f ms  %.2f TFLOP/s' % (err, ms, tf))
