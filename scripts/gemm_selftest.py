import sys; sys.path.insert(0,'.')
from scilmm_b200 import engine as E
for (M,N,K,lower) in [(8192,8192,8192,False),(8192,8192,256,False),(8192,8192,64,False),(16384,16384,256,True),(20000,64,64,False),(140,16000,64,False),(4096,4096,4096,False)]:
    err,ms,tf=E.gemm_selftest(M,N,K,lower=lower,reps=3)
    print(M,N,K,lower,'err %.2e  %.3f ms  %.2f TFLOP/s'%(err,ms,tf), flush=True)
