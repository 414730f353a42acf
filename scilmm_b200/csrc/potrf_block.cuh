// Diagonal-block Cholesky + triangular inverse on FP64 tensor cores (v4).  One CTA (256 threads) per block <= 64 x 64.
//
// The register-blocked predecessor (scripts/potrf_v1_baseline.cuh) spent ~300 instructions per thread and column:
// it was bound by instruction issue, 55 us per block, on the critical chain of every wide front.  Here the
// block is factored 8 columns at a time:
//   (a) warp 0 factors the 8 x 8 diagonal sub-block in registers (lane i = row i, pivots / columns by shuffle) and
//       inverts it (lane c = column c);
//   (b) the rows below take  L21 = A21 Dinv'  as DMMA m8n8k4 tiles (one 8 x 8 tile per warp);
//   (c) the trailing lower triangle takes the rank-8 update  A22 -= L21 L21'  as DMMA tiles.
// The inverse X = L^-1 follows by recursive doubling (block sizes 8, 16, 32), again on DMMA tiles.
// Three barriers per 8 columns, a few hundred instructions per thread in total.
// A non-positive pivot records 1 + global column in *info (smallest failing column wins).
// Replaces the dpotrf CHOLMOD runs on every supernode diagonal block (reference SparseCholesky.py:22-26).
#pragma once
#include "dense_tiles.cuh"

namespace slmm {

constexpr int PQ = 8;                       // sub-block size
constexpr int PLD = NBI + 1;                // shared leading dimension ([row][col])
constexpr int POTRF4_SMEM = 3 * NBI * PLD * 8;

#ifdef POTRF_DEBUG
__device__ long long g_potrf_dbg[8];
#define PDBG(i) if (threadIdx.x == 0 && blockIdx.x == 0) g_potrf_dbg[i] = clock64();
#else
#define PDBG(i)
#endif

// C(8x8) += sign * sum_k A(i,k) B(j,k), k = 0..K-1 (K multiple of 4): A(i,k) = a[i*lda + k], B(j,k) = b[j*ldb + k]
__device__ __forceinline__ void tile_mma(double& c0, double& c1, const double* a, int lda, const double* b, int ldb, int K,
                                         int g, int t) {
  for (int k = 0; k < K; k += 4) dmma884(c0, c1, a[g * lda + k + t], b[g * ldb + k + t]);
}

__global__ void __launch_bounds__(256) potrf_inv_kernel_v4(const PotrfOp* __restrict__ ops, int* __restrict__ info) {
  extern __shared__ double pb_smem[];
  double (*A)[PLD] = reinterpret_cast<double (*)[PLD]>(pb_smem);                  // factor  [row][col]
  double (*X)[PLD] = reinterpret_cast<double (*)[PLD]>(pb_smem + NBI * PLD);      // inverse [row][col]
  double (*T)[PLD] = reinterpret_cast<double (*)[PLD]>(pb_smem + 2 * NBI * PLD);  // products of a doubling level
  __shared__ int fail;
  PDBG(0)
  const PotrfOp op = ops[blockIdx.x];
  const int nb = op.nb, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  if (tid == 0) fail = 0;
  // lower triangle in, identity padding beyond nb (keeps every step well defined), zeros above the diagonal
  {
    double v[NBI * NBI / 256];        // all 16 global loads of a thread in flight before the first shared store
#pragma unroll
    for (int u = 0; u < NBI * NBI / 256; u++) {
      const int q = tid + 256 * u, i = q % NBI, j = q / NBI;
      v[u] = (i == j) ? 1.0 : 0.0;
      if (i < nb && j < nb) v[u] = (i >= j) ? __ldg(op.blk + i + (int64_t)j * op.ld) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < NBI * NBI / 256; u++) {
      const int q = tid + 256 * u, i = q % NBI, j = q / NBI;
      A[i][j] = v[u];
      X[i][j] = 0.0;
    }
  }
  __syncthreads();
  PDBG(1)
  const int nsteps = (nb + PQ - 1) / PQ;
  for (int kq = 0; kq < NBI / PQ; kq++) {
    const int kb = kq * PQ;
    if (kq >= nsteps) {               // identity padding: the inverse of the padding is the identity
      if (tid < PQ) X[kb + tid][kb + tid] = 1.0;
      continue;
    }
    if (warp == 0) {
      // ---- (a) 8 x 8 diagonal sub-block: lane i (0..7) holds row kb+i
      const int i = lane & 7;
      double a[PQ];
#pragma unroll
      for (int c = 0; c < PQ; c++) a[c] = A[kb + i][kb + c];
      int bad = 0;
      double ri[PQ];
#pragma unroll
      for (int j = 0; j < PQ; j++) {
        const double d = __shfl_sync(0xffffffffu, a[j], j);
        if (!(d > 0.0) && bad == 0) bad = kb + j + 1;
        const double rinv = rsqrt(d > 0.0 ? d : 1.0);
        ri[j] = rinv;
        const double l = (i == j) ? d * rinv : a[j] * rinv;
        a[j] = l;
#pragma unroll
        for (int k = j + 1; k < PQ; k++) {
          const double lk = __shfl_sync(0xffffffffu, l, k);
          a[k] = fma(-l, lk, a[k]);
        }
      }
      // inverse of the 8 x 8 factor: lane c (0..7) owns column c; row i of L is fetched from lane i by shuffle
      double x[PQ];
      const int c = lane & 7;
#pragma unroll
      for (int r = 0; r < PQ; r++) {
        double s = (r == c) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < r; k++) {
          const double lrk = __shfl_sync(0xffffffffu, a[k], r);          // L[r][k]
          s = fma(-lrk, (k >= c) ? x[k] : 0.0, s);
        }
        x[r] = (r >= c) ? s * ri[r] : 0.0;                               // 1 / l_rr = rsqrt(pivot)
      }
      if (lane < PQ) {
#pragma unroll
        for (int cc = 0; cc < PQ; cc++) A[kb + i][kb + cc] = (cc <= i) ? a[cc] : 0.0;
#pragma unroll
        for (int r = 0; r < PQ; r++) X[kb + r][kb + c] = x[r];
      }
      if (bad && lane == 0) fail = bad;
    }
    __syncthreads();
    if (fail) break;
    // ---- (b) rows below: L21 = A21 Dinv'   (tile rows kb+8+8w .., K = 8), one tile per warp
    const int r0 = kb + PQ, ntile = (NBI - r0) / PQ;
    if (warp < ntile) {
      const int i0 = r0 + PQ * warp;
      double c0 = 0.0, c1 = 0.0;
      tile_mma(c0, c1, &A[i0][kb], PLD, &X[kb][kb], PLD, PQ, g, t);     // B(j,k) = Dinv(j,k) = X[kb+j][kb+k]
      __syncwarp();
      A[i0 + g][kb + 2 * t] = c0;
      A[i0 + g][kb + 2 * t + 1] = c1;
    }
    __syncthreads();
    // ---- (c) trailing update A22 -= L21 L21' on the lower tiles (ti >= tj)
    const int npair = ntile * (ntile + 1) / 2;
    for (int p = warp; p < npair; p += 8) {
      int ti = 0, rem = p;
      while (rem > ti) { rem -= ti + 1; ti++; }
      const int tj = rem;
      const int i0 = r0 + PQ * ti, j0 = r0 + PQ * tj;
      double c0 = 0.0, c1 = 0.0;
      tile_mma(c0, c1, &A[i0][kb], PLD, &A[j0][kb], PLD, PQ, g, t);
      A[i0 + g][j0 + 2 * t] -= c0;
      A[i0 + g][j0 + 2 * t + 1] -= c1;
    }
    __syncthreads();
  }
  __syncthreads();
  PDBG(2)
  if (fail) {
    if (tid == 0) atomicMin(info, op.colbase + fail);
    return;
  }
  // ---- inverse by recursive doubling over block sizes 8, 16, 32:  X21 = -Binv (C Ainv)
#pragma unroll 1
  for (int m = PQ; m < NBI; m *= 2) {
    const int tpb = (m / PQ) * (m / PQ);          // 8 x 8 tiles per pair block
    const int ntile = (NBI / (2 * m)) * tpb;
    // T = C Ainv : T(i,j) = sum_k C(i,k) Ainv(k,j);  B(j,k) = Ainv(k,j) is read column-wise
    for (int q = warp; q < ntile; q += 8) {
      const int pr = q / tpb, ti = (q % tpb) / (m / PQ), tj = q % (m / PQ);
      const int lo = pr * 2 * m;
      double c0 = 0.0, c1 = 0.0;
      for (int k = 0; k < m; k += 4)
        dmma884(c0, c1, A[lo + m + PQ * ti + g][lo + k + t], X[lo + k + t][lo + PQ * tj + g]);
      T[lo + m + PQ * ti + g][lo + PQ * tj + 2 * t] = c0;
      T[lo + m + PQ * ti + g][lo + PQ * tj + 2 * t + 1] = c1;
    }
    __syncthreads();
    // X21 = -Binv T : X21(i,j) = -sum_k Binv(i,k) T(k,j)
    for (int q = warp; q < ntile; q += 8) {
      const int pr = q / tpb, ti = (q % tpb) / (m / PQ), tj = q % (m / PQ);
      const int lo = pr * 2 * m;
      double c0 = 0.0, c1 = 0.0;
      for (int k = 0; k < m; k += 4)
        dmma884(c0, c1, X[lo + m + PQ * ti + g][lo + m + k + t], T[lo + m + k + t][lo + PQ * tj + g]);
      X[lo + m + PQ * ti + g][lo + PQ * tj + 2 * t] = -c0;
      X[lo + m + PQ * ti + g][lo + PQ * tj + 2 * t + 1] = -c1;
    }
    __syncthreads();
  }
  PDBG(3)
  for (int q = tid; q < NBI * NBI; q += 256) {
    const int i = q % NBI, jj = q / NBI;
    if (i < nb && jj < nb && i >= jj) op.blk[i + (int64_t)jj * op.ld] = A[i][jj];
    if (op.inv_ld == NBI || (i < nb && jj < nb)) op.inv[i + (int64_t)jj * op.inv_ld] = (i < nb && jj < nb) ? X[i][jj] : 0.0;
  }
  PDBG(4)
}

}  // namespace slmm
