// Developer microbenchmark: FP64 tensor-core (DMMA) issue throughput per mma.sync shape on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/dmma_bench.cu -o /tmp/dmma_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE>
__global__ void k(double* out, int iters) {
  double a[8], b[4], c[16][4];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; i++) b[i] = threadIdx.x * 2e-3 + i;
  for (int i = 0; i < 16; i++) for (int j = 0; j < 4; j++) c[i][j] = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      if (SHAPE == 0) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(c[u][0]), "+d"(c[u][1]) : "d"(a[0]), "d"(b[0]));
      } else if (SHAPE == 1) {
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                     : "+d"(c[u][0]), "+d"(c[u][1]), "+d"(c[u][2]), "+d"(c[u][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
      } else if (SHAPE == 2) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+d"(c[u][0]), "+d"(c[u][1]), "+d"(c[u][2]), "+d"(c[u][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
      } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                     : "+d"(c[u][0]), "+d"(c[u][1]), "+d"(c[u][2]), "+d"(c[u][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                       "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
      }
    }
  }
  double s = 0;
  for (int i = 0; i < 16; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE>
void run(const char* name, double flop_per_mma, int warps) {
  double* out;
  cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<SHAPE><<<148, warps * 32>>>(out, 100);
  cudaEventRecord(e0);
  k<SHAPE><<<148, warps * 32>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double flops = 148.0 * warps * iters * 16 * flop_per_mma;
  printf("%-10s warps/SM=%2d  %.2f ms  %.2f TFLOP/s  err=%s\n", name, warps, ms, flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0>("m8n8k4", 2.0 * 8 * 8 * 4, w);
    run<1>("m16n8k4", 2.0 * 16 * 8 * 4, w);
    run<2>("m16n8k8", 2.0 * 16 * 8 * 8, w);
    run<3>("m16n8k16", 2.0 * 16 * 8 * 16, w);
  }
  return 0;
}
