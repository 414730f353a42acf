/* TEST INFRASTRUCTURE ONLY - C helpers of the CPU oracle (oracle/supernodal_cpu.py).
 * Restates, for the CPU baseline, the scatter step CHOLMOD's supernodal factorization performs between its
 * BLAS calls (the library behind sksparse.cholmod.cholesky, reference scilmm/SparseCholesky.py:22-26; CHOLMOD
 * itself is not in /root/reference and not installed here - see oracle/cpu_factor.py).
 * Built by __graft_entry__.build() / oracle/build_oracle.py with: gcc -O3 -shared -fPIC (single-threaded, like the scatter steps between CHOLMOD's BLAS calls). */
#include <stdint.h>

/* child update matrix U (rs x rs column-major, lower triangle valid) is added into the parent front:
 * columns with rel < ns go to the parent's panel (ms x ns, ld = ms), the rest to its update matrix
 * (rsp x rsp, ld = rsp).  Lower triangle only. */
void oracle_extend_add(const double* U, int64_t rs, const int32_t* rel, double* panel, int64_t ms, int64_t ns,
                       double* Up, int64_t rsp) {
  for (int64_t t = 0; t < rs; t++) {
    const int64_t pc = rel[t];
    const double* src = U + t * rs;
    if (pc < ns) {
      double* dst = panel + pc * ms;
      for (int64_t u = t; u < rs; u++) dst[rel[u]] += src[u];
    } else {
      double* dst = Up + (pc - ns) * rsp - ns;
      for (int64_t u = t; u < rs; u++) dst[rel[u]] += src[u];
    }
  }
}

/* x[rows[t], :] -= u[t, :]  (row-major blocks with nrhs columns) */
void oracle_scatter_sub_rows(double* x, const int32_t* rows, int64_t nrows, const double* u, int64_t nrhs) {
  for (int64_t t = 0; t < nrows; t++) {
    double* dst = x + (int64_t)rows[t] * nrhs;
    const double* src = u + t * nrhs;
    for (int64_t j = 0; j < nrhs; j++) dst[j] -= src[j];
  }
}

void oracle_gather_rows(const double* x, const int32_t* rows, int64_t nrows, double* out, int64_t nrhs) {
  for (int64_t t = 0; t < nrows; t++) {
    const double* src = x + (int64_t)rows[t] * nrhs;
    double* dst = out + t * nrhs;
    for (int64_t j = 0; j < nrhs; j++) dst[j] = src[j];
  }
}
