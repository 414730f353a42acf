// Developer microbenchmark: FP64 dependent-chain latency and per-SM throughput on sm_100a (clean, templated, unrolled).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int ILP>
__global__ void k(double* out, long long* cyc, int iters, double y) {
  double x[ILP];
#pragma unroll
  for (int u = 0; u < ILP; u++) x[u] = 1.0 + threadIdx.x * 1e-6 + u * 1e-3;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int rep = 0; rep < 16; rep++) {
#pragma unroll
      for (int u = 0; u < ILP; u++) {
        if (MODE == 0) x[u] = fma(x[u], y, 1e-9);
        else if (MODE == 1) x[u] = rsqrt(x[u]);
        else if (MODE == 2) x[u] = 1.0 / x[u];
        else if (MODE == 3) x[u] = x[u] * y;
        else if (MODE == 4) x[u] = x[u] + y;
        else if (MODE == 5) x[u] = __shfl_sync(0xffffffffu, x[u], (threadIdx.x + 1) & 31);
        else if (MODE == 6) x[u] = sqrt(x[u]);
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int u = 0; u < ILP; u++) s += x[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE, int ILP>
void run(const char* name, int threads) {
  double* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  const int iters = 200;
  k<MODE, ILP><<<1, threads>>>(out, cyc, 10, 0.999999); cudaDeviceSynchronize();
  k<MODE, ILP><<<1, threads>>>(out, cyc, iters, 0.999999); cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = c / (double)(iters * 16);
  printf("%-8s ILP=%d threads=%4d : %.1f cycles per chain link; %.2f warp-instr/cycle/SM\n", name, ILP, threads, per,
         ILP * (threads / 32) / per);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0, 1>("dfma", 32); run<0, 1>("dfma", 256); run<0, 4>("dfma", 256); run<0, 8>("dfma", 1024); run<0, 16>("dfma", 1024);
  run<3, 1>("dmul", 32); run<4, 1>("dadd", 32);
  run<1, 1>("rsqrt", 32); run<1, 4>("rsqrt", 256);
  run<2, 1>("div", 32); run<2, 4>("div", 256);
  run<6, 1>("sqrt", 32);
  run<5, 1>("shfl64", 32);
  return 0;
}
