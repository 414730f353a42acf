// Host-side symbolic analysis for the supernodal LL' factorization of V = sum_k sigma_k A_k.
// Runs once per sparsity pattern (the pattern is invariant across REML iterations); replaces the
// per-call CHOLMOD analyze the reference triggers through sksparse (reference scilmm/SparseCholesky.py:22-26,
// called from :92 on every likelihood evaluation).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace slmm {

// Leading dimension of a panel with `nrow` rows: rounded up to even, so that every panel column starts on a 16-byte
// boundary - the alignment bulk tensor copies (TMA) need for the global strides of a tensor map.
inline int64_t panel_ld(int64_t nrow) { return (nrow + 1) & ~(int64_t)1; }

enum Ordering { ORD_NATURAL = 0, ORD_GIVEN = 1, ORD_METIS = 2, ORD_MINDEG = 3,
                ORD_METIS_FAST = 4 /* one separator per bisection instead of the best of three */ };

struct SymbolicOptions {
  int ordering = ORD_METIS;
  // relaxed supernode amalgamation (same shape as CHOLMOD's nrelax/zrelax rule; small supernodes merge more
  // eagerly than CHOLMOD's {4,16,48} because a launch costs more than a few padded columns, but large ones only
  // when the merge adds < 1 % explicit zeros: at 0.08 the padding in the top fronts cost 9 % of the factorization
  // time at the 250K config, measured)
  int nrelax[3] = {8, 32, 96};
  double zrelax[3] = {0.8, 0.2, 0.01};
  int max_super_cols = 1 << 30;   // cap on columns per supernode (0 = none)
};

struct Symbolic {
  int n = 0;
  std::vector<int32_t> perm;    // perm[new] = old   (factor.P() of the Factor protocol)
  std::vector<int32_t> iperm;   // iperm[old] = new
  std::vector<int32_t> parent;  // column elimination tree in the new order
  std::vector<int32_t> colcount;

  int nsuper = 0;
  std::vector<int32_t> sn_first;    // [nsuper+1] first column of each supernode
  std::vector<int32_t> sn_nrow;     // rows in the panel (ns own columns first, then the rows below)
  std::vector<int32_t> sn_parent;   // supernodal etree (-1 root)
  std::vector<int32_t> sn_depth;    // 0 at roots
  std::vector<int64_t> sn_rowptr;   // [nsuper+1] into rows[]
  std::vector<int64_t> sn_lptr;     // [nsuper+1] into the panel storage (column-major, ld = panel_ld(sn_nrow))
  std::vector<int32_t> rows;        // concatenated sorted row lists
  std::vector<int32_t> rel;         // aligned with rows[]: position of a below-row inside the parent's row list
  std::vector<int32_t> col2sn;      // [n]
  std::vector<int32_t> child_ptr, child_idx;   // children of each supernode (ascending)

  int nlevels = 0;                  // depth levels; level d holds supernodes with sn_depth == d
  std::vector<int32_t> level_ptr, level_sn;

  int64_t nnzL = 0;       // sum of column counts (entries of L including the diagonal)
  double flops = 0;       // sum of colcount^2  (CHOLMOD 'fl' convention)
  int64_t lsize = 0;      // doubles in the panel storage (>= nnzL because of the dense trapezoids)
  int ncomponents = 0;
  int max_front_rows = 0, max_super_cols = 0;
  double t_order = 0, t_symbolic = 0;
};

// pattern: full symmetric CSR/CSC (both triangles; diagonal optional), int32, n x n.
// user_perm (perm[new]=old) is used when opt.ordering == ORD_GIVEN.
void analyze(int n, const int32_t* indptr, const int32_t* indices, const int32_t* user_perm,
             const SymbolicOptions& opt, Symbolic& out);

// position map of original matrix entries into the panel storage.
// (rowidx_of_entry, colidx_of_entry) are given as a CSR/CSC pattern (symmetric => orientation is immaterial);
// target[e] = offset into panel storage for the copy with new_row >= new_col, -1 for the mirrored copy.
// tri != 0 selects CHOLMOD's one-triangle semantics (see the definition).
// Entries outside the analysed pattern raise std::runtime_error.
void entry_map(const Symbolic& S, const int32_t* indptr, const int32_t* indices, int64_t* target, int tri = 0);

}  // namespace slmm
