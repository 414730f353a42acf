// Dense tile engine for the supernodal factorization and the supernode-blocked triangular solves.
// Every dense operation of the numeric phase is one of two work-item kinds, executed from device-resident
// operation lists that the host builds once per sparsity pattern:
//   * PotrfOp  - Cholesky of one diagonal block (<= 64 x 64) in shared memory + its triangular inverse
//   * GemmOp   - C (+/-)= A * B^T on FP64 tensor cores (DMMA mma.sync m8n8k4), generic element strides so
//                that the same kernel serves panel updates, Schur complements (lower-masked SYRK), the
//                TRSM-by-inverse steps and the multi-RHS solve updates (row-major RHS blocks, gathered rows).
// These replace the dpotrf/dtrsm/dsyrk/dgemm calls CHOLMOD's supernodal kernels make on CPU BLAS for the
// reference (sksparse.cholmod.cholesky, reference scilmm/SparseCholesky.py:22-26) and CHOLMOD's solve_A
// (factor(b), reference :30,32,52,100).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace slmm {

constexpr int NBI = 64;   // diagonal block size (POTRF / inverse granularity)
constexpr int POTRF_SMEM = 2 * NBI * (NBI + 1) * 8;

enum GemmFlags { GF_LOWER = 1, GF_ACCUM = 2, GF_NEG = 4 };

struct GemmOp {
  double* C;
  const double* A;
  const double* B;
  const int32_t* a_kidx;   // optional gather: A(i,k) = A[i*a_si + a_kidx[k]*a_sk]
  int64_t c_si, c_sj, a_si, a_sk, b_sj, b_sk;   // element strides: C(i,j), A(i,k), B(j,k)
  int32_t M, N, K;
  int32_t flags;
  int32_t tile_start, tiles_m, tiles_n, pad;
};

struct PotrfOp {
  double* blk;      // diagonal block inside the panel (column-major)
  double* inv;      // NBI x NBI column-major slot for the inverse of the factored block
  int32_t ld, nb;
  int32_t colbase;  // global (permuted) index of the first column, for failure reporting
  int32_t pad;
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------------------
// GEMM tiles: one CTA per TM x TN tile of C.  K is consumed in slabs of KS through a two-stage shared
// memory ring (register-staged prefetch).  Shared tiles are stored [k][m] with a +4 pad so that the DMMA
// fragment loads (8 rows x 4 k per quad layout) are bank-conflict free.
template <int TM, int TN, int NWM, int NWN>
__global__ void __launch_bounds__(32 * NWM * NWN) gemm_tiles_kernel(const GemmOp* __restrict__ ops, int nops) {
  constexpr int NT = 32 * NWM * NWN;
  constexpr int KS = 16;
  constexpr int WM = TM / NWM, WN = TN / NWN;
  constexpr int MI = WM / 8, NI = WN / 8;
  constexpr int LDA = TM + 4, LDB = TN + 4;
  constexpr int EA = TM * KS / NT, EB = TN * KS / NT;
  static_assert(TM * KS % NT == 0 && TN * KS % NT == 0, "tile/threads mismatch");
  extern __shared__ double smem[];
  double* As = smem;                      // [2][KS][LDA]
  double* Bs = smem + 2 * KS * LDA;       // [2][KS][LDB]

  // locate the operation this tile belongs to
  int lo = 0, hi = nops - 1;
  const int tile = blockIdx.x;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (ops[mid].tile_start <= tile) lo = mid; else hi = mid - 1;
  }
  const GemmOp op = ops[lo];
  const int local = tile - op.tile_start;
  const int tm = local % op.tiles_m, tn = local / op.tiles_m;
  const int tm0 = tm * TM, tn0 = tn * TN;
  if ((op.flags & GF_LOWER) && tm0 + TM <= tn0) return;   // tile entirely above the diagonal

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp % NWM) * WM, wn0 = (warp / NWM) * WN;
  const int Mrem = op.M - tm0, Nrem = op.N - tn0;
  const bool a_kcontig = (op.a_sk == 1 && op.a_kidx == nullptr && op.a_si != 1);
  const bool b_kcontig = (op.b_sk == 1 && op.b_sj != 1);
  const double* Abase = op.A + (int64_t)tm0 * op.a_si;
  const double* Bbase = op.B + (int64_t)tn0 * op.b_sj;

  double acc[MI][NI][2];
#pragma unroll
  for (int mi = 0; mi < MI; mi++)
#pragma unroll
    for (int ni = 0; ni < NI; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

  double ra[EA], rb[EB];
  auto load_slab = [&](int k0) {
#pragma unroll
    for (int e = 0; e < EA; e++) {
      const int q = e * NT + tid;
      int i, k;
      if (a_kcontig) { k = q % KS; i = q / KS; } else { i = q % TM; k = q / TM; }
      double v = 0.0;
      if (i < Mrem && k0 + k < op.K) {
        const int64_t kk = op.a_kidx ? (int64_t)op.a_kidx[k0 + k] : (int64_t)(k0 + k);
        v = Abase[(int64_t)i * op.a_si + kk * op.a_sk];
      }
      ra[e] = v;
    }
#pragma unroll
    for (int e = 0; e < EB; e++) {
      const int q = e * NT + tid;
      int j, k;
      if (b_kcontig) { k = q % KS; j = q / KS; } else { j = q % TN; k = q / TN; }
      double v = 0.0;
      if (j < Nrem && k0 + k < op.K) v = Bbase[(int64_t)j * op.b_sj + (int64_t)(k0 + k) * op.b_sk];
      rb[e] = v;
    }
  };
  auto store_slab = [&](int stage) {
    double* as = As + stage * KS * LDA;
    double* bs = Bs + stage * KS * LDB;
#pragma unroll
    for (int e = 0; e < EA; e++) {
      const int q = e * NT + tid;
      int i, k;
      if (a_kcontig) { k = q % KS; i = q / KS; } else { i = q % TM; k = q / TM; }
      as[k * LDA + i] = ra[e];
    }
#pragma unroll
    for (int e = 0; e < EB; e++) {
      const int q = e * NT + tid;
      int j, k;
      if (b_kcontig) { k = q % KS; j = q / KS; } else { j = q % TN; k = q / TN; }
      bs[k * LDB + j] = rb[e];
    }
  };

  const int nslab = (op.K + KS - 1) / KS;
  if (nslab > 0) {
    load_slab(0);
    store_slab(0);
  }
  __syncthreads();
  for (int s = 0; s < nslab; s++) {
    if (s + 1 < nslab) load_slab((s + 1) * KS);
    const double* as = As + (s & 1) * KS * LDA + wm0 + g;
    const double* bs = Bs + (s & 1) * KS * LDB + wn0 + g;
#pragma unroll
    for (int kk = 0; kk < KS; kk += 4) {
      double af[MI], bf[NI];
#pragma unroll
      for (int mi = 0; mi < MI; mi++) af[mi] = as[(kk + t) * LDA + mi * 8];
#pragma unroll
      for (int ni = 0; ni < NI; ni++) bf[ni] = bs[(kk + t) * LDB + ni * 8];
#pragma unroll
      for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
    }
    if (s + 1 < nslab) store_slab((s + 1) & 1);
    __syncthreads();
  }

  // epilogue: C = [C] +/- acc, optionally only on/below the diagonal of the region
  const bool accum = op.flags & GF_ACCUM, neg = op.flags & GF_NEG, lower = op.flags & GF_LOWER;
#pragma unroll
  for (int mi = 0; mi < MI; mi++) {
    const int i = tm0 + wm0 + mi * 8 + g;
    if (i >= op.M) continue;
#pragma unroll
    for (int ni = 0; ni < NI; ni++) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int j = tn0 + wn0 + ni * 8 + 2 * t + h;
        if (j >= op.N || (lower && i < j)) continue;
        double* cp = op.C + (int64_t)i * op.c_si + (int64_t)j * op.c_sj;
        double v = neg ? -acc[mi][ni][h] : acc[mi][ni][h];
        if (accum) v += *cp;
        *cp = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Diagonal-block Cholesky + inverse.  One CTA (256 threads) per block; the block lives in shared memory.
// A non-positive pivot records 1 + global column in *info (smallest failing column wins).
__global__ void __launch_bounds__(256) potrf_inv_kernel(const PotrfOp* __restrict__ ops, int* __restrict__ info) {
  extern __shared__ double potrf_smem[];
  double (*S)[NBI + 1] = reinterpret_cast<double (*)[NBI + 1]>(potrf_smem);
  double (*X)[NBI + 1] = reinterpret_cast<double (*)[NBI + 1]>(potrf_smem + NBI * (NBI + 1));
  __shared__ int fail;
  const PotrfOp op = ops[blockIdx.x];
  const int nb = op.nb, tid = threadIdx.x;
  if (tid == 0) fail = 0;
  for (int q = tid; q < NBI * NBI; q += 256) {
    const int i = q % NBI, j = q / NBI;
    S[i][j] = (i < nb && j < nb && i >= j) ? op.blk[i + (int64_t)j * op.ld] : 0.0;
    X[i][j] = 0.0;
  }
  __syncthreads();
  for (int j = 0; j < nb; j++) {
    const double d = S[j][j];
    if (!(d > 0.0)) {
      if (tid == 0) { fail = 1; atomicMin(info, op.colbase + j + 1); }
      break;                                   // uniform: every thread reads the same S[j][j]
    }
    const double r = sqrt(d);
    __syncthreads();
    if (tid == 0) S[j][j] = r;
    for (int i = j + 1 + tid; i < nb; i += 256) S[i][j] /= r;
    __syncthreads();
    // trailing update of the lower triangle: S[i][k] -= S[i][j] * S[k][j], j < k <= i
    const int m = nb - j - 1;
    for (int q = tid; q < m * m; q += 256) {
      const int i = j + 1 + q % m, k = j + 1 + q / m;
      if (k <= i) S[i][k] -= S[i][j] * S[k][j];
    }
    __syncthreads();
  }
  __syncthreads();
  if (fail) return;
  // inverse of the lower-triangular block, one column per thread, zero-padded so every thread walks the
  // same k range (broadcast reads of S, conflict-free reads of X)
  if (tid < nb) {
    const int c = tid;
    for (int i = 0; i < nb; i++) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = 0; k < i; k++) s -= S[i][k] * X[k][c];
      X[i][c] = (i >= c) ? s / S[i][i] : 0.0;
    }
  }
  __syncthreads();
  for (int q = tid; q < NBI * NBI; q += 256) {
    const int i = q % NBI, j = q / NBI;
    if (i < nb && j < nb && i >= j) op.blk[i + (int64_t)j * op.ld] = S[i][j];
    op.inv[i + j * NBI] = (i < nb && j < nb) ? X[i][j] : 0.0;
  }
}

}  // namespace slmm
