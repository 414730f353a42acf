// Developer harness: diagonal-block Cholesky + inverse kernels (v1 register-blocked vs v4 DMMA-blocked),
// correctness against a host reference and latency (1 block) / throughput (many blocks).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I scilmm_b200/csrc -I scripts scripts/potrf_bench.cu -o scripts/potrf_bench.bin
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define POTRF_DEBUG
#include "potrf_block.cuh"
#include "potrf_v1_baseline.cuh"
using namespace slmm;

static void host_chol(std::vector<double>& a, int n, int ld) {
  for (int j = 0; j < n; j++) {
    for (int k = 0; k < j; k++) for (int i = j; i < n; i++) a[i + j * ld] -= a[i + k * ld] * a[j + k * ld];
    double d = std::sqrt(a[j + j * ld]);
    for (int i = j; i < n; i++) a[i + j * ld] /= d;
  }
}

int main() {
  const int ld = 70, nblk = 2048;
  for (int nb : {64, 37, 16, 5}) {
    std::vector<double> h((size_t)nblk * ld * 64), ref;
    srand(1);
    for (int b = 0; b < nblk; b++) {
      std::vector<double> g(64 * 64);
      for (auto& v : g) v = rand() / (double)RAND_MAX - 0.5;
      for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) {
        double s = (i == j) ? 4.0 : 0.0;
        for (int k = 0; k < 64; k++) s += g[i * 64 + k] * g[j * 64 + k] / 16.0;
        if (i < ld) h[(size_t)b * ld * 64 + i + j * ld] = s;
      }
    }
    ref = h;
    { std::vector<double> blk(ref.begin(), ref.begin() + ld * 64); host_chol(blk, nb, ld); std::copy(blk.begin(), blk.end(), ref.begin()); }
    double *d_a, *d_inv; PotrfOp* d_ops; int* d_info;
    cudaMalloc(&d_a, h.size() * 8); cudaMalloc(&d_inv, (size_t)nblk * 4096 * 8); cudaMalloc(&d_ops, nblk * sizeof(PotrfOp)); cudaMalloc(&d_info, 4);
    std::vector<PotrfOp> ops(nblk);
    for (int b = 0; b < nblk; b++) ops[b] = {d_a + (size_t)b * ld * 64, d_inv + (size_t)b * 4096, ld, nb, b * 64, 64};
    cudaMemcpy(d_ops, ops.data(), nblk * sizeof(PotrfOp), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(potrf_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRF_SMEM);
    cudaFuncSetAttribute(potrf_inv_kernel_v4, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRF4_SMEM);
    for (int ver = 1; ver <= 2; ver++) {
      for (int grid : {1, nblk}) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e9;
        for (int rep = 0; rep < 5; rep++) {
          cudaMemcpy(d_a, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
          int big = 0x7fffffff; cudaMemcpy(d_info, &big, 4, cudaMemcpyHostToDevice);
          cudaEventRecord(e0);
          if (ver == 1) potrf_inv_kernel<<<grid, 256, POTRF_SMEM>>>(d_ops, d_info);
          else potrf_inv_kernel_v4<<<grid, 256, POTRF4_SMEM>>>(d_ops, d_info);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
        }
        std::vector<double> out(ld * 64), inv(4096);
        cudaMemcpy(out.data(), d_a, ld * 64 * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(inv.data(), d_inv, 4096 * 8, cudaMemcpyDeviceToHost);
        double err = 0, ierr = 0;
        for (int j = 0; j < nb; j++) for (int i = j; i < nb; i++) err = std::fmax(err, std::fabs(out[i + j * ld] - ref[i + j * ld]));
        for (int i = 0; i < nb; i++) for (int j = 0; j < nb; j++) {     // L * X = I
          double s = 0; for (int k = 0; k < nb; k++) s += (k <= i ? out[i + k * ld] : 0.0) * inv[k + j * 64];
          ierr = std::fmax(ierr, std::fabs(s - (i == j ? 1.0 : 0.0)));
        }
        if (ver == 2 && grid == 1) { long long dbg[8]; cudaMemcpyFromSymbol(dbg, g_potrf_dbg, sizeof(dbg));
          printf("   cycles: load %lld  factor %lld  inverse %lld  store %lld\n", dbg[1]-dbg[0], dbg[2]-dbg[1], dbg[3]-dbg[2], dbg[4]-dbg[3]); }
        int info; cudaMemcpy(&info, d_info, 4, cudaMemcpyDeviceToHost);
        printf("nb=%2d v%d grid=%4d  %.1f us  L err %.2e  L*inv-I %.2e  info %s  %s\n", nb, ver, grid, best * 1e3, err, ierr,
               info == 0x7fffffff ? "ok" : "FAIL", cudaGetErrorString(cudaGetLastError()));
      }
    }
    // not-positive-definite detection
    h[3 + 3 * ld] = -1.0;
    cudaMemcpy(d_a, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    int big = 0x7fffffff; cudaMemcpy(d_info, &big, 4, cudaMemcpyHostToDevice);
    potrf_inv_kernel_v4<<<1, 256, POTRF4_SMEM>>>(d_ops, d_info);
    int info; cudaMemcpy(&info, d_info, 4, cudaMemcpyDeviceToHost);
    printf("nb=%2d v2 non-PD at column 3 -> info %d (expect 4)\n", nb, info);
    cudaFree(d_a); cudaFree(d_inv); cudaFree(d_ops); cudaFree(d_info);
  }
  return 0;
}
