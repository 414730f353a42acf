/* TEST INFRASTRUCTURE ONLY - C helpers of the CPU oracle (oracle/supernodal_cpu.py).
 * Restates, for the CPU baseline, the scatter step CHOLMOD's supernodal factorization performs between its
 * BLAS calls (the library behind sksparse.cholmod.cholesky, reference scilmm/SparseCholesky.py:22-26; CHOLMOD
 * itself is not in /root/reference and not installed here - see oracle/cpu_factor.py).
 * Built by __graft_entry__.build() / oracle/build_oracle.py with: gcc -O3 -shared -fPIC (single-threaded, like the scatter steps between CHOLMOD's BLAS calls). */
#include <stdint.h>

/* child update matrix U (rs x rs column-major, lower triangle valid) is added into the parent front:
 * columns with rel < ns go to the parent's panel (ns columns, leading dimension ms - the engine pads it to an even
 * number of rows), the rest to its update matrix
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

/* ---------------------------------------------------------------------------------------------------------
 * Independent symbolic analysis for the parity tests (oracle/symbolic_ref.py): given a symmetric pattern that is
 * ALREADY permuted (CSR, both triangles, sorted rows), the elimination tree, the column counts of L, fundamental
 * supernodes and their row structures - what CHOLMOD's analyze step computes for the reference
 * (sksparse.cholmod.cholesky, reference scilmm/SparseCholesky.py:22-26).  Deliberately the textbook algorithms
 * (Liu's etree with path compression; counts and structures by walking every row subtree, O(nnz(L))), sharing no
 * code with the product's csrc/symbolic.cpp (Gilbert-Ng-Peyton counts, relaxed amalgamation), so that nnz(L),
 * the column counts and the factor itself are cross-checked without the product library. */
#include <stdlib.h>

void oracle_etree(int64_t n, const int32_t* ap, const int32_t* ai, int32_t* parent) {
  int32_t* anc = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  for (int64_t j = 0; j < n; j++) {
    parent[j] = -1;
    anc[j] = -1;
    for (int32_t p = ap[j]; p < ap[j + 1]; p++) {
      int32_t i = ai[p];
      while (i != -1 && i < j) {
        const int32_t next = anc[i];
        anc[i] = (int32_t)j;
        if (next == -1) parent[i] = (int32_t)j;
        i = next;
      }
    }
  }
  free(anc);
}

/* colcount[j] = number of entries of column j of L, diagonal included */
void oracle_colcounts(int64_t n, const int32_t* ap, const int32_t* ai, const int32_t* parent, int32_t* colcount) {
  int32_t* mark = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  for (int64_t j = 0; j < n; j++) { mark[j] = -1; colcount[j] = 1; }
  for (int64_t i = 0; i < n; i++) {
    mark[i] = (int32_t)i;
    for (int32_t p = ap[i]; p < ap[i + 1]; p++) {
      int32_t j = ai[p];
      if (j >= i) continue;
      while (mark[j] != i) {          /* climb the row subtree of i */
        mark[j] = (int32_t)i;
        colcount[j]++;
        j = parent[j];
      }
    }
  }
  free(mark);
}

/* Row structures of the supernodes' first columns.  first_of[j] = supernode id when j is the first column of a
 * supernode, else -1; rowptr[s] = where the list of supernode s starts in rows[] (its length is colcount of the
 * first column); fill[] is scratch of nsuper int64.  Rows arrive in ascending order, so the lists are sorted. */
void oracle_supernode_rows(int64_t n, const int32_t* ap, const int32_t* ai, const int32_t* parent,
                           const int32_t* first_of, const int64_t* rowptr, int64_t nsuper, int32_t* rows) {
  int32_t* mark = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nsuper > 0 ? nsuper : 1));
  for (int64_t s = 0; s < nsuper; s++) fill[s] = rowptr[s];
  for (int64_t j = 0; j < n; j++) mark[j] = -1;
  for (int64_t i = 0; i < n; i++) {
    mark[i] = (int32_t)i;
    if (first_of[i] >= 0) rows[fill[first_of[i]]++] = (int32_t)i;      /* the diagonal entry */
    for (int32_t p = ap[i]; p < ap[i + 1]; p++) {
      int32_t j = ai[p];
      if (j >= i) continue;
      while (mark[j] != i) {
        mark[j] = (int32_t)i;
        if (first_of[j] >= 0) rows[fill[first_of[j]]++] = (int32_t)i;
        j = parent[j];
      }
    }
  }
  free(mark);
  free(fill);
}

/* target[e] = offset of permuted entry (max, min) of original entry e inside the panel storage, -1 for the copy
 * that lands above the diagonal (both triangles are stored).  ap/ai: ORIGINAL order CSR; iperm[old] = new. */
int32_t oracle_entry_map(int64_t n, const int32_t* ap, const int32_t* ai, const int32_t* iperm, const int32_t* col2sn,
                         const int32_t* sn_first, const int32_t* sn_nrow, const int64_t* sn_rowptr,
                         const int64_t* sn_lptr, const int32_t* rows, int64_t* target) {
  for (int64_t r = 0; r < n; r++) {
    const int32_t ir = iperm[r];
    for (int32_t p = ap[r]; p < ap[r + 1]; p++) {
      const int32_t ic = iperm[ai[p]];
      if (ic > ir) { target[p] = -1; continue; }
      const int32_t s = col2sn[ic];
      const int32_t* rb = rows + sn_rowptr[s];
      int64_t lo = 0, hi = sn_nrow[s];
      while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (rb[mid] < ir) lo = mid + 1; else hi = mid; }
      if (lo >= sn_nrow[s] || rb[lo] != ir) return 1;
      target[p] = sn_lptr[s] + (int64_t)(ic - sn_first[s]) * sn_nrow[s] + lo;
    }
  }
  return 0;
}
