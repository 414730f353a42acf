// Developer microbenchmark: dependent-chain latency of DFMA / rsqrt / LDS+DFMA / barrier on sm_100a (1 warp and 8 warps).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int mode, int iters) {
  __shared__ double sh[512];
  sh[threadIdx.x] = 1.0 + threadIdx.x * 1e-3;
  __syncthreads();
  double x = 1.0 + threadIdx.x * 1e-6, y = 0.999999;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    if (mode == 0) { x = fma(x, y, 1e-9); }
    else if (mode == 1) { x = rsqrt(x) + 1.0; }
    else if (mode == 2) { x = fma(sh[(threadIdx.x + (int)x) & 255], y, x * 1e-9); }
    else if (mode == 3) { x = fma(x, y, 1e-9); __syncthreads(); }
    else if (mode == 4) { x = sqrt(x) + 1.0; }
    else if (mode == 5) { x = 1.0 / x + 1.0; }
    else if (mode == 6) { float f = (float)x; f = fmaf(f, 0.999f, 1e-6f); x = (double)f; }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  const char* names[] = {"dfma chain", "rsqrt+add chain", "lds+dfma chain", "dfma+barrier", "sqrt+add", "div+add", "cvt+ffma+cvt"};
  for (int threads : {32, 256})
    for (int m = 0; m < 7; m++) {
      k<<<1, threads>>>(out, cyc, m, 1000); cudaDeviceSynchronize();
      k<<<1, threads>>>(out, cyc, m, 10000); cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("threads=%3d %-18s %.1f cycles/iter\n", threads, names[m], c / 10000.0);
    }
  return 0;
}
