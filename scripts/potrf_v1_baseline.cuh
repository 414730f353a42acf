// Developer baseline for scripts/potrf_bench.cu: the register-blocked diagonal-block kernel the engine used before
// the DMMA-based one (scilmm_b200/csrc/potrf_block.cuh).  Not part of the library.
#pragma once
#include "dense_tiles.cuh"

namespace slmm {

constexpr int POTRF_SMEM = 2 * NBI * (NBI + 1) * 8;

// ---------------------------------------------------------------------------------------------------------
// Diagonal-block Cholesky + inverse.  One CTA (256 threads) per block <= 64 x 64.
//   * The block lives in REGISTERS: thread (bi = tid % 16, bj = tid / 16) owns the 4 x 4 sub-block
//     rows 4bi.., cols 4bj.. (right-looking updates are 16 register FMAs, no shared-memory read-modify-write).
//   * Per column j only the current (still unscaled) column is exchanged through shared memory, double
//     buffered, so the loop needs ONE barrier per column.  Every thread derives 1/l_jj = rsqrt(a_jj) itself.
//   * The finished columns of L are mirrored to shared memory (Lf) and row j of the inverse,
//     X[j,:] = (e_j - L[j,:j] X[:j,:]) / l_jj, is computed in the tail of the same iteration by thread
//     (quarter = tid % 4, col = tid / 4) with a quad shuffle reduction.
// A non-positive pivot records 1 + global column in *info (smallest failing column wins).
__global__ void __launch_bounds__(256) potrf_inv_kernel(const PotrfOp* __restrict__ ops, int* __restrict__ info) {
  extern __shared__ double potrf_smem[];
  double (*Lf)[NBI + 1] = reinterpret_cast<double (*)[NBI + 1]>(potrf_smem);
  double (*X)[NBI + 1] = reinterpret_cast<double (*)[NBI + 1]>(potrf_smem + NBI * (NBI + 1));
  __shared__ double colbuf[2][NBI];
  const PotrfOp op = ops[blockIdx.x];
  const int nb = op.nb, tid = threadIdx.x;
  const int bi = tid & 15, bj = tid >> 4;
  const int r0 = 4 * bi, c0 = 4 * bj;
  double a[4][4];
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int row = r0 + r, col = c0 + c;
      a[r][c] = (row < nb && col < nb && row >= col) ? op.blk[row + (int64_t)col * op.ld] : 0.0;
    }
  for (int q = tid; q < NBI * (NBI + 1); q += 256) { (&Lf[0][0])[q] = 0.0; (&X[0][0])[q] = 0.0; }
  if (bj == 0) {
#pragma unroll
    for (int r = 0; r < 4; r++) colbuf[0][r0 + r] = a[r][0];
  }
  __syncthreads();
  const int qtr = tid & 3, xc = tid >> 2;
  double rinv_prev = 0.0;
  // Row j-1 of the inverse is computed in the head of iteration j, in the same basic block as the rsqrt chain of
  // column j: FP64 latency (not throughput) bounds this kernel, and the two dependency chains are independent.
  auto inverse_row = [&](int jr, double rinv_r) {
    double s0 = 0.0, s1 = 0.0;
    int k = xc + qtr;
#pragma unroll 4
    for (; k + 4 < jr; k += 8) {
      s0 += Lf[jr][k] * X[k][xc];
      s1 += Lf[jr][k + 4] * X[k + 4][xc];
    }
    if (k < jr) s0 += Lf[jr][k] * X[k][xc];
    double sum = s0 + s1;
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    if (qtr == 0 && xc <= jr) X[jr][xc] = ((jr == xc ? 1.0 : 0.0) - sum) * rinv_r;
  };
  for (int j = 0; j < nb; j++) {
    const double* cb = colbuf[j & 1];
    const double d = cb[j];
    if (!(d > 0.0)) {                            // uniform: every thread reads the same value
      if (tid == 0) atomicMin(info, op.colbase + j + 1);
      return;
    }
    const double rinv = rsqrt(d);                // l_jj = d * rsqrt(d)
    if (j > 0) inverse_row(j - 1, rinv_prev);
    rinv_prev = rinv;
    double li[4], lk[4];
#pragma unroll
    for (int r = 0; r < 4; r++) li[r] = cb[r0 + r] * rinv;
#pragma unroll
    for (int c = 0; c < 4; c++) lk[c] = cb[c0 + c] * rinv;
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
      for (int r = 0; r < 4; r++)
        if (c0 + c > j && r0 + r >= c0 + c) a[r][c] -= li[r] * lk[c];
    const int jb = j >> 2;
    if (bj == jb) {                              // owners of column j: publish the finished column of L
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const int row = r0 + r;
        Lf[row][j] = row > j ? li[r] : (row == j ? d * rinv : 0.0);
      }
    }
    if (j + 1 < nb && bj == ((j + 1) >> 2)) {    // owners of column j+1: publish it (updated, unscaled)
      const int cn = (j + 1) & 3;
#pragma unroll
      for (int r = 0; r < 4; r++) {
        double v = a[r][0];
#pragma unroll
        for (int c = 1; c < 4; c++) v = (c == cn) ? a[r][c] : v;
        colbuf[(j + 1) & 1][r0 + r] = v;
      }
    }
    __syncthreads();
  }
  if (nb > 0) inverse_row(nb - 1, rinv_prev);
  __syncthreads();
  for (int q = tid; q < NBI * NBI; q += 256) {
    const int i = q % NBI, jj = q / NBI;
    if (i < nb && jj < nb && i >= jj) op.blk[i + (int64_t)jj * op.ld] = Lf[i][jj];
    // a private 64 x 64 slot is padded with zeros; inside a wider block inverse only the nb x nb part exists
    if (op.inv_ld == NBI || (i < nb && jj < nb)) op.inv[i + (int64_t)jj * op.inv_ld] = (i < nb && jj < nb) ? X[i][jj] : 0.0;
  }
}

}  // namespace slmm
