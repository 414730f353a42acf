// Narrow right-hand sides (nrhs <= 16): the supernode-blocked triangular solves and L*Z as HBM-STREAMING kernels.
//
// factor(b) for the c+1 fixed-effect columns, a single phenotype vector, the K and K(K+1)/2 columns of compute_hess
// (reference scilmm/SparseCholesky.py:30,32,100,149,153) and the per-GPU probe slice at 8 GPUs all have at most 16
// columns.  There every GemmOp of the solve schedules is  C(M x N) (+/-)= A(M x K) B(N x K)^T  with M = nrhs: 2*M flops
// per 8 bytes of L - far below the machine balance - so the bound is HBM (each sweep reads L once: 8*nnz(L) bytes),
// not the tensor pipe.  The DMMA tile kernel wastes a 64-row tile on <= 16 rows, pays a producer/consumer hand-off per
// 16-wide K slab and reached 1.2 TB/s on the 12-column solve of the 250K config; these kernels keep the M accumulators
// of one output column in registers, stage the small A operand in shared memory and stream B with many independent
// coalesced loads per thread.  Same GemmOp descriptors, same schedules, same deterministic split-K reduction.
//
// Two flavours, by the memory layout of B (the factor panel or an inverse block):
//   F1  B(j,k) = B[j + k*b_sk]   consecutive OUTPUT columns j are contiguous (forward sweep, L*Z):
//       thread <-> output column, loop over k; a warp reads 32 consecutive doubles per k.
//   F2  B(j,k) = B[j*b_sj + k]   consecutive k are contiguous (backward sweep: L' and gathered rows):
//       warp <-> output column, lanes stride k; one shuffle reduction per column.
#pragma once
#include "dense_tiles.cuh"

namespace slmm {

constexpr int SK_F1_COLS = 128;    // output columns per CTA, flavour 1 (= threads)
constexpr int SK_F1_KC = 256;      // k chunk staged in shared memory per pass
constexpr int SK_F2_COLS = 64;     // output columns per CTA, flavour 2 (8 warps x 8 columns)
constexpr int SK_F2_KMAX = 512;    // an F2 work item holds its whole K range of A in shared memory

template <int MT>
__global__ void __launch_bounds__(SK_F1_COLS) skinny_f1_kernel(const GemmOp* __restrict__ ops,
                                                               const int32_t* __restrict__ tile_op) {
  __shared__ double As[SK_F1_KC * MT];                 // [k][i]
  const int tile = blockIdx.x;
  const GemmOp& op = ops[tile_op[tile]];
  const int j = (tile - op.tile_start) * SK_F1_COLS + threadIdx.x;
  const int M = op.M, N = op.N, K = op.K, flags = op.flags;
  const bool active = j < N;
  // lower-triangular B (L11 of L*Z): B(j,k) = 0 for k > j; op.pad = first k of this (split) part
  const int kend = (flags & GF_TRIL_B) ? min(K, max(0, min(N - 1, (tile - op.tile_start) * SK_F1_COLS + SK_F1_COLS - 1) + 1 - op.pad)) : K;
  const int kmine = (flags & GF_TRIL_B) ? min(K, max(0, j + 1 - op.pad)) : K;
  const double* __restrict__ Bj = op.B + j;
  const int64_t b_sk = op.b_sk;
  double acc[MT];
#pragma unroll
  for (int i = 0; i < MT; i++) acc[i] = 0.0;
  for (int k0 = 0; k0 < kend; k0 += SK_F1_KC) {
    const int kc = min(SK_F1_KC, kend - k0);
    __syncthreads();
    for (int q = threadIdx.x; q < kc * MT; q += SK_F1_COLS) {
      const int k = q / MT, i = q - k * MT;
      As[q] = i < M ? op.A[(int64_t)i + (int64_t)(k0 + k) * op.a_sk] : 0.0;
    }
    __syncthreads();
    if (!active) continue;
    const int kl = min(kc, kmine - k0);
    int k = 0;
    for (; k + 8 <= kl; k += 8) {                      // 8 independent loads in flight per thread
      double b[8];
#pragma unroll
      for (int u = 0; u < 8; u++) b[u] = __ldcs(Bj + (int64_t)(k0 + k + u) * b_sk);
#pragma unroll
      for (int u = 0; u < 8; u++)
#pragma unroll
        for (int i = 0; i < MT; i++) acc[i] += As[(k + u) * MT + i] * b[u];
    }
    for (; k < kl; k++) {
      const double b = __ldcs(Bj + (int64_t)(k0 + k) * b_sk);
#pragma unroll
      for (int i = 0; i < MT; i++) acc[i] += As[k * MT + i] * b;
    }
  }
  if (!active) return;
  double* c = op.C + (int64_t)j * op.c_sj;
  const bool accum = flags & GF_ACCUM, neg = flags & GF_NEG;
#pragma unroll
  for (int i = 0; i < MT; i++)
    if (i < M) {
      const double v = neg ? -acc[i] : acc[i];
      c[i] = accum ? c[i] + v : v;
    }
}

template <int MT>
__global__ void __launch_bounds__(256) skinny_f2_kernel(const GemmOp* __restrict__ ops,
                                                        const int32_t* __restrict__ tile_op) {
  extern __shared__ double As2[];                      // [i][K] (lanes read consecutive k: conflict free)
  const int tile = blockIdx.x;
  const GemmOp& op = ops[tile_op[tile]];
  const int j0 = (tile - op.tile_start) * SK_F2_COLS;
  const int M = op.M, N = op.N, K = op.K, flags = op.flags;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int32_t* __restrict__ kidx = op.a_kidx;
  for (int q = threadIdx.x; q < K * MT; q += 256) {
    const int i = q / K, k = q - i * K;
    double v = 0.0;
    if (i < M) v = op.A[(int64_t)i + (int64_t)(kidx ? kidx[k] : k) * op.a_sk];
    As2[q] = v;
  }
  __syncthreads();
  const bool accum = flags & GF_ACCUM, neg = flags & GF_NEG;
  for (int jj = warp; jj < SK_F2_COLS; jj += 8) {
    const int j = j0 + jj;
    if (j >= N) break;
    const double* __restrict__ Bj = op.B + (int64_t)j * op.b_sj;
    double acc[MT];
#pragma unroll
    for (int i = 0; i < MT; i++) acc[i] = 0.0;
    int k = lane;
    for (; k + 96 < K; k += 128) {                     // 4 independent coalesced loads in flight per lane
      double b[4];
#pragma unroll
      for (int u = 0; u < 4; u++) b[u] = __ldcs(Bj + k + 32 * u);
#pragma unroll
      for (int u = 0; u < 4; u++)
#pragma unroll
        for (int i = 0; i < MT; i++) acc[i] += As2[i * K + k + 32 * u] * b[u];
    }
    for (; k < K; k += 32) {
      const double b = __ldcs(Bj + k);
#pragma unroll
      for (int i = 0; i < MT; i++) acc[i] += As2[i * K + k] * b;
    }
    // reduce the MT accumulators over the warp; lane i ends up with row i
    double mine = 0.0;
#pragma unroll
    for (int i = 0; i < MT; i++) {
      double v = acc[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == i) mine = v;
    }
    if (lane < M) {
      double* c = op.C + lane + (int64_t)j * op.c_sj;
      const double v = neg ? -mine : mine;
      *c = accum ? *c + v : v;
    }
  }
}

}  // namespace slmm
