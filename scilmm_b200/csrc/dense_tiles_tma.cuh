// FP64 DMMA tile GEMM with TMA-staged operands (cp.async.bulk.tensor, SASS UTMALDG).
//
// Same contract as gemm_tiles_kernel<128,128,...> (dense_tiles.cuh): C (+/-)= A B^T for one 128 x 128 tile per CTA,
// GemmOp descriptors, lower-masked / accumulating / negated variants - for the operations whose two operands are
// column-major blocks with unit row stride and an even column stride (panel updates, Schur complements, the
// triangular panel multiplies: > 95 % of the factorization's flops).  What changes is how a K slab reaches shared
// memory: one elected producer thread issues 16 bulk tensor copies per slab (8 boxes of 16 rows x 16 k per operand)
// against a tensor map of the operand, the copies land with the hardware 128-byte swizzle and complete on the
// stage's "full" mbarrier through its transaction count; no thread computes addresses or issues 8-byte copies.
//
// Shared-memory layout of one operand slab: 8 boxes, each [k = 0..15][16 doubles] = 16 rows of 128 bytes; the
// 16-byte chunk c of row k sits at chunk c ^ (k & 7) (CU_TENSOR_MAP_SWIZZLE_128B).  DMMA m8n8k4 wants, per lane
// (g = lane / 4, t = lane % 4), A[row g][k = kk + t]: with rows taken in natural order two of the four k of a
// half-warp fall on the same banks.  The kernel therefore maps the 8 logical fragment rows of a 16-row box to the
// physical rows {0,1,8,9,2,3,10,11} (+4 for the second 8-row group): within a half-warp the chunks are {c, c^4} and
// the XOR with k only permutes the low two chunk bits - conflict free.  The same permutation is applied to the
// fragment columns (B) and undone in the epilogue's C indexing, so it is invisible outside the kernel.
//
// Alignment: a bulk tensor copy must start on a 16-byte boundary (an odd element offset in the contiguous dimension
// raises an illegal-instruction fault), but panel blocks start at row ns or c1, which may be odd.  The tensor map of
// such an operand is based one element EARLIER (aligned) and the op's tile grid is shifted with it: tile row i'
// stands for row i' - 1 of the op (op.pad bit 0 for A, bit 1 for B), the extra leading row / column is masked in the
// epilogue.  Shared-memory addressing is the same for aligned and shifted operands.
#pragma once
#include <cuda.h>

#include "dense_tiles.cuh"

namespace slmm {

constexpr int TMA_KS = 16;                 // K slab
constexpr int TMA_STAGES = 6;
constexpr int TMA_BOX_BYTES = TMA_KS * 128;                     // one box: 16 k-rows of 16 doubles
constexpr int TMA_OPERAND_BYTES = 8 * TMA_BOX_BYTES;            // 128 tile rows
constexpr int TMA_STAGE_BYTES = 2 * TMA_OPERAND_BYTES;
constexpr int TMA_SMEM = TMA_STAGES * TMA_STAGE_BYTES + 1024 + 2 * TMA_STAGES * 8;   // + alignment slack + barriers
constexpr int TMA_THREADS = 32 * (8 + NPW);

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(smem_dst),
               "l"(map), "r"(c0), "r"(c1), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}

__global__ void __launch_bounds__(TMA_THREADS, 1) gemm_tiles_tma_kernel(const GemmOp* __restrict__ ops,
                                                                         const int32_t* __restrict__ tile_op,
                                                                         const CUtensorMap* __restrict__ maps) {
  constexpr int TM = 128, TN = 128, NWM = 2, NWN = 4, NCW = 8;
  constexpr int KS = TMA_KS, STAGES = TMA_STAGES;
  constexpr int WM = TM / NWM, WN = TN / NWN;      // 64 x 32 per warp
  constexpr int MI = WM / 8, NI = WN / 8;
  extern __shared__ uint8_t tma_smem_raw[];
  const uint32_t raw = (uint32_t)__cvta_generic_to_shared(tma_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                  // swizzle-128B boxes need 1024-byte alignment
  uint8_t* gen = tma_smem_raw + (base - raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(gen + STAGES * TMA_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;

  const int tile = blockIdx.x;
  const int opi = tile_op[tile];
  const GemmOp& op = ops[opi];
  const int local = tile - op.tile_start;
  const int tiles_m = op.tiles_m;
  const int tm0 = (local % tiles_m) * TM, tn0 = (local / tiles_m) * TN;
  const int flags = op.flags;
  const int sa = op.pad & 1, sb = (op.pad >> 1) & 1;           // tile rows / columns are offset by one element
  if ((flags & GF_LOWER) && tm0 + TM - 1 - sa < tn0 - sb) return;   // tile entirely above the diagonal

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = op.M, N = op.N, K = op.K;
  const int nslab = (K + KS - 1) / KS;
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp >= NCW) {
    // ------------------------------------------------------------------ producer warpgroup (one lane issues TMA)
    reg_dealloc<56>();
    const int pw = warp - NCW;
    if ((flags & GF_ACCUM) && op.c_si == 1) {                    // pull the C tile into L2 under the main loop
      const int pt = pw * 32 + lane;
      const int j = tn0 + pt - sb, i0 = max(0, tm0 - sa);
      if (pt < TN && j >= 0 && j < N) {
        const double* ccol = op.C + (int64_t)i0 + (int64_t)j * op.c_sj;
        const int rows = min(TM, M - i0);
        const int r_first = (flags & GF_LOWER) ? max(0, j - i0) : 0;
        for (int r = (r_first / 16) * 16; r < rows; r += 16) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(ccol + r));
      }
    }
    if (pw == 0 && lane == 0) {
      const CUtensorMap* mapA = maps + 2 * (int64_t)opi;
      const CUtensorMap* mapB = mapA + 1;
      const int a_r = tm0, b_r = tn0;       // always even: the maps of odd-offset operands are based one element early
      for (int s = 0; s < nslab; s++) {
        const int stage = s % STAGES, use = s / STAGES;
        if (use > 0) mbar_wait(empty_bar + stage, (use - 1) & 1);
        mbar_expect_tx(full_bar + stage, TMA_STAGE_BYTES);
        const uint32_t dst_a = base + stage * TMA_STAGE_BYTES, dst_b = dst_a + TMA_OPERAND_BYTES;
        const int k0 = s * KS;
#pragma unroll
        for (int b = 0; b < 8; b++) tma_load_2d(dst_a + b * TMA_BOX_BYTES, mapA, a_r + 16 * b, k0, full_bar + stage);
#pragma unroll
        for (int b = 0; b < 8; b++) tma_load_2d(dst_b + b * TMA_BOX_BYTES, mapB, b_r + 16 * b, k0, full_bar + stage);
      }
    }
    return;
  }
  reg_alloc<224>();

  // -------------------------------------------------------------------- compute warps
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp % NWM) * WM, wn0 = (warp / NWM) * WN;
  // physical row of logical fragment row q (0..7) inside a 16-row box: {0,1,8,9,2,3,10,11}
  const int colg = (g & 1) + ((g >> 1) & 1) * 8 + (g >> 2) * 2;
  const int chg = colg >> 1;                                     // 16-byte chunk of that row (second 8-row group: +2)
  // byte offset inside a box for (row group parity p, k parity q = (kk >> 2) & 1), without the k-row term
  int xo[2][2];
#pragma unroll
  for (int p = 0; p < 2; p++)
#pragma unroll
    for (int q = 0; q < 2; q++) xo[p][q] = (((chg + 2 * p) ^ (4 * q + t)) << 4) + (g & 1) * 8 + t * 128;
  double acc[MI][NI][2];
#pragma unroll
  for (int mi = 0; mi < MI; mi++)
#pragma unroll
    for (int ni = 0; ni < NI; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

  for (int s = 0; s < nslab; s++) {
    const int stage = s % STAGES;
    mbar_wait(full_bar + stage, (s / STAGES) & 1);
    const uint8_t* as = gen + stage * TMA_STAGE_BYTES + (wm0 >> 4) * TMA_BOX_BYTES;
    const uint8_t* bs = gen + stage * TMA_STAGE_BYTES + TMA_OPERAND_BYTES + (wn0 >> 4) * TMA_BOX_BYTES;
#pragma unroll
    for (int kk = 0; kk < KS; kk += 4) {
      double af[MI], bf[NI];
#pragma unroll
      for (int mi = 0; mi < MI; mi++)
        af[mi] = *reinterpret_cast<const double*>(as + (mi >> 1) * TMA_BOX_BYTES + kk * 128 + xo[mi & 1][(kk >> 2) & 1]);
#pragma unroll
      for (int ni = 0; ni < NI; ni++)
        bf[ni] = *reinterpret_cast<const double*>(bs + (ni >> 1) * TMA_BOX_BYTES + kk * 128 + xo[ni & 1][(kk >> 2) & 1]);
#pragma unroll
      for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar + stage);
  }

  // epilogue: logical fragment (mi, g) -> tile row, (ni, 2t+h) -> tile column through the same permutation
  const bool accum = flags & GF_ACCUM, neg = flags & GF_NEG, lower = flags & GF_LOWER;
  double* Cb = op.C;
  const int64_t c_si = op.c_si, c_sj = op.c_sj;
#pragma unroll
  for (int mi = 0; mi < MI; mi++) {
    const int i = tm0 + wm0 + 16 * (mi >> 1) + 4 * (mi & 1) + colg - sa;
    double old[NI][2];
    bool ok[NI][2];
    int jj[NI][2];
#pragma unroll
    for (int ni = 0; ni < NI; ni++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int q = 2 * t + h;                                  // logical column 0..7 of the 8-column group
        const int j = tn0 + wn0 + 16 * (ni >> 1) + 4 * (ni & 1) + (q & 1) + ((q >> 1) & 1) * 8 + (q >> 2) * 2 - sb;
        jj[ni][h] = j;
        ok[ni][h] = (i >= 0) && (i < M) && (j >= 0) && (j < N) && !(lower && i < j);
        old[ni][h] = 0.0;
      }
    if (accum) {
#pragma unroll
      for (int ni = 0; ni < NI; ni++)
#pragma unroll
        for (int h = 0; h < 2; h++)
          if (ok[ni][h]) old[ni][h] = Cb[(int64_t)i * c_si + (int64_t)jj[ni][h] * c_sj];
    }
#pragma unroll
    for (int ni = 0; ni < NI; ni++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const double a = acc[mi][ni][h];
        if (ok[ni][h]) Cb[(int64_t)i * c_si + (int64_t)jj[ni][h] * c_sj] = old[ni][h] + (neg ? -a : a);
      }
  }
}

}  // namespace slmm
