// Dense tile engine for the supernodal factorization and the supernode-blocked triangular solves.
// Every dense operation of the numeric phase is one of two work-item kinds, executed from device-resident
// operation lists that the host builds once per sparsity pattern:
//   * PotrfOp  - Cholesky of one diagonal block (<= 64 x 64) + its triangular inverse (kernel: potrf_block.cuh)
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

enum GemmFlags { GF_LOWER = 1, GF_ACCUM = 2, GF_NEG = 4, GF_WS = 8 /* C is an offset into the split-K workspace */,
                 GF_BIGTILE = 16 /* host-side only: keep the 128 x 128 tile configuration */,
                 GF_TRIL_B = 32 /* B(j,k) = 0 for k > j (lower-triangular B): a tile stops at k = tn0 + TN */,
                 GF_WS2 = 64 /* like GF_WS, in the second workspace (bulk-stream ops of the solves) */ };

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

// Second half of a deterministic split-K GEMM: C (+/-)= sum over the S partial products parked in the workspace.
struct ReduceOp {
  double* C;
  const double* ws;        // S consecutive M x N column-major partials
  int64_t c_si, c_sj;
  int32_t M, N, S, flags;
  int32_t block_start, pad;
};

struct PotrfOp {
  double* blk;      // diagonal block inside the panel (column-major)
  double* inv;      // NBI x NBI column-major slot for the inverse of the factored block
  int32_t ld, nb;
  int32_t colbase;  // global (permuted) index of the first column, for failure reporting
  int32_t inv_ld;   // leading dimension of the inverse slot (NBI for the private slots, wider inside a block inverse)
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------------------
// GEMM tiles: one CTA per TM x TN tile of C, warp-specialised.
//   * NWM*NWN compute warps: LDS fragment loads + DMMA only (their instruction stream stays tensor-dense);
//   * one producer warp: streams K in slabs of KS=16 into a STAGES-deep shared-memory ring with 8-byte cp.async
//     (LDGSTS: no register staging, zero-fill for ragged edges, any leading dimension, gathered rows);
//   * full/empty mbarriers per stage (cp.async.mbarrier.arrive.noinc on the producer side), no __syncthreads in
//     the main loop, so compute warps never wait for each other's address arithmetic.
// Shared tiles are stored [k][m] with a +4 pad: the DMMA fragment loads (8 rows x 4 k per quad) are conflict free.
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, int src_bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, int parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
  }
}

constexpr int NPW = 4;   // producer warps (one warpgroup, so that setmaxnreg can hand its registers to the compute warps)

template <int REGS>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(REGS)); }
template <int REGS>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(REGS)); }

// One operand of one K slab -> shared memory, the share of producer warp `pw`.
//   rows-major mode (unit or generic stride along the tile dimension): the warp copies k rows pw*4 .. pw*4+3 of the
//     slab, 32 consecutive tile rows per LDGSTS (one coalesced 256-byte request when the stride is 1);
//   k-contiguous mode: 16 consecutive k of 2 tile rows per LDGSTS; the warp copies tile rows [pw*T/4, (pw+1)*T/4).
template <int T, int LD>
__device__ __noinline__ void produce_operand(double* __restrict__ sm, const double* __restrict__ base, int64_t s_i,
                                                int64_t s_k, const int32_t* __restrict__ kidx, bool kcontig, int rem,
                                                int k0, int K, int pw, int lane) {
  constexpr int KS = 16;
  if (!kcontig) {
    const int64_t lane_off = (int64_t)lane * s_i, chunk = 32 * s_i;
#pragma unroll
    for (int kq = 0; kq < KS / NPW; kq++) {
      const int k = pw * (KS / NPW) + kq;
      const bool kok = k0 + k < K;
      int64_t gk = k0 + k;
      if (kidx != nullptr) gk = kok ? kidx[k0 + k] : 0;
      const double* src = base + gk * s_k + lane_off;
      double* dst = sm + k * LD + lane;
#pragma unroll
      for (int c = 0; c < T / 32; c++) {
        const bool ok = kok && (lane + 32 * c < rem);
        cp_async8(dst + 32 * c, ok ? src + c * chunk : base, ok ? 8 : 0);
      }
    }
  } else {
    const int k = lane & 15, r = lane >> 4;
    const bool kok = k0 + k < K;
    const double* src = base + (k0 + k);
    double* dst = sm + k * LD;
#pragma unroll
    for (int c = 0; c < T / 2 / NPW; c++) {
      const int i = 2 * (pw * (T / 2 / NPW) + c) + r;
      const bool ok = kok && i < rem;
      cp_async8(dst + i, ok ? src + (int64_t)i * s_i : base, ok ? 8 : 0);
    }
  }
}

template <int TM, int TN, int NWM, int NWN, int STAGES>
__global__ void __launch_bounds__(32 * (NWM * NWN + NPW), 1) gemm_tiles_kernel(const GemmOp* __restrict__ ops,
                                                                                const int32_t* __restrict__ tile_op) {
  constexpr int NCW = NWM * NWN;           // compute warps (a multiple of 4: whole warpgroups)
  constexpr int KS = 16;
  constexpr int WM = TM / NWM, WN = TN / NWN;
  constexpr int MI = WM / 8, NI = WN / 8;
  constexpr int LDA = TM + 4, LDB = TN + 4;
  // register rebalancing only where the compute warps need it (8 x 4 accumulator fragments per thread)
  constexpr bool REBALANCE = (NCW == 8);
  static_assert(TM % 32 == 0 && TN % 32 == 0 && NCW % 4 == 0, "producer mapping / warpgroup layout");
  extern __shared__ double smem[];
  double* As = smem;                           // [STAGES][KS][LDA]
  double* Bs = smem + STAGES * KS * LDA;       // [STAGES][KS][LDB]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(Bs + STAGES * KS * LDB);   // [STAGES]
  uint64_t* empty_bar = full_bar + STAGES;                                       // [STAGES]

  // the operation this tile belongs to: one table lookup (a binary search over tile_start costs ~15 dependent
  // L2 round trips per CTA, which dominated launches made of thousands of tiny tiles)
  const int tile = blockIdx.x;
  const GemmOp& op = ops[tile_op[tile]];
  const int local = tile - op.tile_start;
  const int tiles_m = op.tiles_m;
  const int tm0 = (local % tiles_m) * TM, tn0 = (local / tiles_m) * TN;
  const int flags = op.flags;
  if ((flags & GF_LOWER) && tm0 + TM <= tn0) return;   // tile entirely above the diagonal (whole CTA exits)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = op.M, N = op.N;
  // triangular B: the rest of K multiplies zeros; op.pad = first k of a split part (0 for whole ops)
  const int K = (flags & GF_TRIL_B) ? min(op.K, max(0, tn0 + TN - op.pad)) : op.K;
  const int nslab = (K + KS - 1) / KS;
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(full_bar + s, 32 * NPW); mbar_init(empty_bar + s, NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp >= NCW) {
    // ------------------------------------------------------------------ producer warpgroup
    if (REBALANCE) reg_dealloc<56>();
    const int pw = warp - NCW;
    const int Mrem = M - tm0, Nrem = N - tn0;
    const int64_t a_si = op.a_si, a_sk = op.a_sk, b_sj = op.b_sj, b_sk = op.b_sk;
    const int32_t* kidx = op.a_kidx;
    const bool a_kcontig = (a_sk == 1 && kidx == nullptr && a_si != 1);
    const bool b_kcontig = (b_sk == 1 && b_sj != 1);
    const double* Abase = op.A + (int64_t)tm0 * a_si;
    const double* Bbase = op.B + (int64_t)tn0 * b_sj;
    // Accumulating tiles read C in the epilogue: pull the tile into L2 now, under the main loop, so that the
    // epilogue's dependent loads cost an L2 hit instead of a DRAM round trip (one CTA per SM: nothing else would
    // hide them).  Column-major C only (unit row stride): one 128-byte line per 16 rows.
    if ((flags & GF_ACCUM) && op.c_si == 1) {
      const int pt = pw * 32 + lane;                       // 0..127: one tile column per producer thread
      if (pt < TN && tn0 + pt < N) {
        const double* ccol = op.C + (int64_t)tm0 + (int64_t)(tn0 + pt) * op.c_sj;
        const int rows = min(TM, M - tm0);
        const int r_first = (flags & GF_LOWER) ? max(0, tn0 + pt - tm0) : 0;
        for (int r = (r_first / 16) * 16; r < rows; r += 16)
          asm volatile("prefetch.global.L2 [%0];\n" ::"l"(ccol + r));
      }
    }
    for (int s = 0; s < nslab; s++) {
      const int stage = s % STAGES, use = s / STAGES;
      if (use > 0) mbar_wait(empty_bar + stage, (use - 1) & 1);
      const int k0 = s * KS;
      produce_operand<TM, LDA>(As + stage * KS * LDA, Abase, a_si, a_sk, kidx, a_kcontig, Mrem, k0, K, pw, lane);
      produce_operand<TN, LDB>(Bs + stage * KS * LDB, Bbase, b_sj, b_sk, nullptr, b_kcontig, Nrem, k0, K, pw, lane);
      mbar_cp_async_arrive(full_bar + stage);    // arrives (once per lane) when this lane's copies have landed
    }
    return;
  }
  if (REBALANCE) reg_alloc<224>();

  // -------------------------------------------------------------------- compute warps
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp % NWM) * WM, wn0 = (warp / NWM) * WN;
  double acc[MI][NI][2];
#pragma unroll
  for (int mi = 0; mi < MI; mi++)
#pragma unroll
    for (int ni = 0; ni < NI; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

  for (int s = 0; s < nslab; s++) {
    const int stage = s % STAGES;
    mbar_wait(full_bar + stage, (s / STAGES) & 1);
    const double* as = As + stage * KS * LDA + wm0 + g;
    const double* bs = Bs + stage * KS * LDB + wn0 + g;
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
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar + stage);     // this warp no longer reads the stage
  }

  // epilogue: C = [C] +/- acc, optionally only on/below the diagonal of the region.  The read-modify-write is
  // done in batches of 2*NI*2 independent loads per thread (a load-use chain per element would serialise on
  // the memory latency and cost more than the K loop for K = 64).
  const bool accum = flags & GF_ACCUM, neg = flags & GF_NEG, lower = flags & GF_LOWER;
  double* Cb = op.C;
  const int64_t c_si = op.c_si, c_sj = op.c_sj;
  constexpr int MB = 1;
#pragma unroll
  for (int mb = 0; mb < MI; mb += MB) {
    double old[MB][NI][2];
    bool ok[MB][NI][2];
#pragma unroll
    for (int m2 = 0; m2 < MB; m2++) {
      const int i = tm0 + wm0 + (mb + m2) * 8 + g;
#pragma unroll
      for (int ni = 0; ni < NI; ni++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int j = tn0 + wn0 + ni * 8 + 2 * t + h;
          ok[m2][ni][h] = (i < M) && (j < N) && !(lower && i < j);
          old[m2][ni][h] = 0.0;
        }
    }
    if (accum) {
#pragma unroll
      for (int m2 = 0; m2 < MB; m2++) {
        const int i = tm0 + wm0 + (mb + m2) * 8 + g;
#pragma unroll
        for (int ni = 0; ni < NI; ni++)
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int j = tn0 + wn0 + ni * 8 + 2 * t + h;
            if (ok[m2][ni][h]) old[m2][ni][h] = Cb[(int64_t)i * c_si + (int64_t)j * c_sj];
          }
      }
    }
#pragma unroll
    for (int m2 = 0; m2 < MB; m2++) {
      const int i = tm0 + wm0 + (mb + m2) * 8 + g;
#pragma unroll
      for (int ni = 0; ni < NI; ni++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int j = tn0 + wn0 + ni * 8 + 2 * t + h;
          const double a = acc[mb + m2][ni][h];
          if (ok[m2][ni][h]) Cb[(int64_t)i * c_si + (int64_t)j * c_sj] = old[m2][ni][h] + (neg ? -a : a);
        }
    }
  }
}

// fixed-order reduction of split-K partials (no atomics: results are bit-reproducible run to run)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const ReduceOp* __restrict__ ops, int nops) {
  int lo = 0, hi = nops - 1;
  const int blk = blockIdx.x;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (ops[mid].block_start <= blk) lo = mid; else hi = mid - 1;
  }
  const ReduceOp op = ops[lo];
  const int64_t total = (int64_t)op.M * op.N;
  const int64_t e = (int64_t)(blk - op.block_start) * 256 + threadIdx.x;
  if (e >= total) return;
  double sum = 0.0;
  for (int s = 0; s < op.S; s++) sum += op.ws[(int64_t)s * total + e];
  const int i = (int)(e % op.M), j = (int)(e / op.M);
  double* cp = op.C + (int64_t)i * op.c_si + (int64_t)j * op.c_sj;
  const double v = (op.flags & GF_NEG) ? -sum : sum;
  *cp = (op.flags & GF_ACCUM) ? *cp + v : v;
}

}  // namespace slmm
