// Diagonal-block Cholesky + triangular inverse (v3).  One CTA (256 threads) per block <= 64 x 64.
//
// On B200 a DEPENDENT FP64 operation costs ~120 cycles (DFMA chain 119, rsqrt 177, divide 171 cycles per link,
// measured with scripts/fp64_latency.cu), so this kernel is bound by the length of its dependency chains, not by
// throughput.  Both halves are organised to keep those chains short:
//   * factorization: right-looking, the block in REGISTERS (thread (bi,bj) owns a 4 x 4 sub-block), one barrier
//     per column.  The update uses a -= (a_rj a_cj) * (1/d): the products do not wait for the pivot, and the
//     reciprocal runs beside the rsqrt that scales the finished column, so a column costs
//     LDS -> divide -> DFMA -> STS -> barrier (~400 cycles) instead of LDS -> rsqrt -> DMUL -> DFMA -> ...
//   * inverse: recursive doubling  inv([A 0; C B]) = [A^-1 0; -B^-1 C A^-1  B^-1]  over block sizes 1,2,...,32,
//     every dot product split over independent accumulators: depth ~50 links instead of ~64 * 8.
// A non-positive pivot records 1 + global column in *info (smallest failing column wins).
// Replaces the dpotrf CHOLMOD runs on every supernode diagonal block (reference SparseCholesky.py:22-26).
#pragma once
#include "dense_tiles.cuh"

namespace slmm {

constexpr int PLD = NBI + 1;                // shared leading dimension
constexpr int POTRF3_SMEM = 3 * NBI * PLD * 8;

#ifdef POTRF_DEBUG
__device__ long long g_potrf_dbg[8];
#define PDBG(i) if (threadIdx.x == 0 && blockIdx.x == 0) g_potrf_dbg[i] = clock64();
#else
#define PDBG(i)
#endif

__global__ void __launch_bounds__(256) potrf_inv_kernel_v3(const PotrfOp* __restrict__ ops, int* __restrict__ info) {
  extern __shared__ double pb_smem[];
  double (*Lf)[PLD] = reinterpret_cast<double (*)[PLD]>(pb_smem);                  // factor [row][col]
  double (*X)[PLD] = reinterpret_cast<double (*)[PLD]>(pb_smem + NBI * PLD);       // inverse
  double (*Tm)[PLD] = reinterpret_cast<double (*)[PLD]>(pb_smem + 2 * NBI * PLD);  // C * A^-1 of the current level
  __shared__ double colbuf[2][NBI];
  PDBG(0)
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
      // identity padding beyond nb keeps every later step well defined
      a[r][c] = (row < nb && col < nb) ? (row >= col ? op.blk[row + (int64_t)col * op.ld] : 0.0) : (row == col ? 1.0 : 0.0);
    }
  for (int q = tid; q < NBI * PLD; q += 256) { (&Lf[0][0])[q] = 0.0; (&X[0][0])[q] = 0.0; }
  if (bj == 0) {
#pragma unroll
    for (int r = 0; r < 4; r++) colbuf[0][r0 + r] = a[r][0];
  }
  __syncthreads();
  PDBG(1)
  for (int j = 0; j < NBI; j++) {
    const double* cb = colbuf[j & 1];
    const double d = cb[j];
    if (!(d > 0.0)) {                            // uniform: every thread reads the same value
      if (tid == 0) atomicMin(info, op.colbase + j + 1);
      return;
    }
    const double rd = 1.0 / d;                   // on the critical path (feeds the update of column j+1)
    double cr[4], cc[4];
#pragma unroll
    for (int r = 0; r < 4; r++) cr[r] = cb[r0 + r];
#pragma unroll
    for (int c = 0; c < 4; c++) cc[c] = cb[c0 + c];
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
      for (int r = 0; r < 4; r++)
        if (c0 + c > j && r0 + r >= c0 + c) a[r][c] = fma(-(cr[r] * cc[c]), rd, a[r][c]);
    if (j + 1 < NBI && bj == ((j + 1) >> 2)) {   // owners of column j+1: publish it (updated, unscaled)
      const int cn = (j + 1) & 3;
#pragma unroll
      for (int r = 0; r < 4; r++) {
        double v = a[r][0];
#pragma unroll
        for (int c = 1; c < 4; c++) v = (c == cn) ? a[r][c] : v;
        colbuf[(j + 1) & 1][r0 + r] = v;
      }
    }
    if (bj == (j >> 2)) {                        // owners of column j: scale and store the finished column (off chain)
      const double rinv = rsqrt(d);
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const int row = r0 + r;
        Lf[row][j] = row > j ? cr[r] * rinv : (row == j ? d * rinv : 0.0);
      }
      if (bi == (j >> 2)) X[j][j] = rinv;        // 1 / l_jj: level 0 of the inverse
    }
    __syncthreads();
  }
  PDBG(2)
  // ---- inverse by recursive doubling: blocks of size m -> 2m
#pragma unroll 1
  for (int m = 1; m < NBI; m *= 2) {
    const int npair = NBI / (2 * m), per = m * m;
    // T = C * A^-1 :  T[r][c] = sum_{k=c}^{m-1} C[r][k] Ainv[k][c]
    for (int q = tid; q < npair * per; q += 256) {
      const int p = q / per, r = (q / m) % m, c = q % m;
      const int lo = p * 2 * m;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int k = c;
      for (; k + 3 < m; k += 4) {
        s0 = fma(Lf[lo + m + r][lo + k], X[lo + k][lo + c], s0);
        s1 = fma(Lf[lo + m + r][lo + k + 1], X[lo + k + 1][lo + c], s1);
        s2 = fma(Lf[lo + m + r][lo + k + 2], X[lo + k + 2][lo + c], s2);
        s3 = fma(Lf[lo + m + r][lo + k + 3], X[lo + k + 3][lo + c], s3);
      }
      for (; k < m; k++) s0 = fma(Lf[lo + m + r][lo + k], X[lo + k][lo + c], s0);
      Tm[lo + m + r][lo + c] = (s0 + s1) + (s2 + s3);
    }
    __syncthreads();
    // X21 = -B^-1 * T :  X[lo+m+r][lo+c] = -sum_{k=0}^{r} Binv[r][k] T[k][c]
    for (int q = tid; q < npair * per; q += 256) {
      const int p = q / per, r = (q / m) % m, c = q % m;
      const int lo = p * 2 * m;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int k = 0;
      for (; k + 3 <= r; k += 4) {
        s0 = fma(X[lo + m + r][lo + m + k], Tm[lo + m + k][lo + c], s0);
        s1 = fma(X[lo + m + r][lo + m + k + 1], Tm[lo + m + k + 1][lo + c], s1);
        s2 = fma(X[lo + m + r][lo + m + k + 2], Tm[lo + m + k + 2][lo + c], s2);
        s3 = fma(X[lo + m + r][lo + m + k + 3], Tm[lo + m + k + 3][lo + c], s3);
      }
      for (; k <= r; k++) s0 = fma(X[lo + m + r][lo + m + k], Tm[lo + m + k][lo + c], s0);
      X[lo + m + r][lo + c] = -((s0 + s1) + (s2 + s3));
    }
    __syncthreads();
  }
  PDBG(3)
  for (int q = tid; q < NBI * NBI; q += 256) {
    const int i = q % NBI, jj = q / NBI;
    if (i < nb && jj < nb && i >= jj) op.blk[i + (int64_t)jj * op.ld] = Lf[i][jj];
    op.inv[i + (int64_t)jj * op.inv_ld] = (i < nb && jj < nb) ? X[i][jj] : 0.0;
  }
  PDBG(4)
}

}  // namespace slmm
