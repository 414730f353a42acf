// Host symbolic analysis: components -> fill-reducing ordering -> elimination tree -> postorder ->
// column counts -> relaxed supernodes -> supernodal row structures -> depth levels -> extend-add maps.
// Plays the role CHOLMOD's analyze step plays for the reference (scilmm/SparseCholesky.py:22-26), but
// runs once per pattern instead of once per likelihood evaluation.
#include "symbolic.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <numeric>
#include <cstdio>
#include <stdexcept>
#include <thread>

extern "C" int METIS_NodeND(int64_t* nvtxs, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* options,
                            int64_t* perm, int64_t* iperm);
extern "C" int METIS_SetDefaultOptions(int64_t* options);

namespace slmm {
namespace {

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---- exact minimum degree on a small component (dense bitset adjacency) --------------------------------
void small_mindeg(int m, const std::vector<std::vector<int>>& adj, std::vector<int>& order) {
  const int W = (m + 63) / 64;
  std::vector<uint64_t> bits((size_t)m * W, 0);
  auto row = [&](int i) { return bits.data() + (size_t)i * W; };
  for (int i = 0; i < m; i++)
    for (int j : adj[i])
      if (j != i) row(i)[j >> 6] |= 1ull << (j & 63);
  std::vector<char> done(m, 0);
  order.clear();
  order.reserve(m);
  std::vector<int> nb;
  for (int step = 0; step < m; step++) {
    int best = -1, bestdeg = 1 << 30;
    for (int i = 0; i < m; i++) {
      if (done[i]) continue;
      int d = 0;
      for (int w = 0; w < W; w++) d += __builtin_popcountll(row(i)[w]);
      if (d < bestdeg) { bestdeg = d; best = i; }
    }
    done[best] = 1;
    order.push_back(best);
    nb.clear();
    for (int w = 0; w < W; w++) {
      uint64_t x = row(best)[w];
      while (x) { int b = __builtin_ctzll(x); nb.push_back(w * 64 + b); x &= x - 1; }
    }
    for (int u : nb) {                       // neighbours become a clique, pivot disappears
      uint64_t* ru = row(u);
      for (int w = 0; w < W; w++) ru[w] |= row(best)[w];
      ru[u >> 6] &= ~(1ull << (u & 63));
      ru[best >> 6] &= ~(1ull << (best & 63));
    }
    std::fill(row(best), row(best) + W, 0ull);
  }
}

static int host_threads() {
  int T = (int)std::thread::hardware_concurrency();
  if (const char* e = getenv("SLMM_HOST_THREADS")) T = atoi(e);
  return std::max(1, std::min(T, 32));
}

// ---- ordering -----------------------------------------------------------------------------------------
void compute_ordering(int n, const int32_t* ap, const int32_t* ai, int method, std::vector<int32_t>& perm,
                      int& ncomp) {
  perm.resize(n);
  const bool fast_nd = method == ORD_METIS_FAST;
  if (fast_nd) method = ORD_METIS;
  if (method == ORD_NATURAL) {
    std::iota(perm.begin(), perm.end(), 0);
    ncomp = 0;
    return;
  }
  // METIS option overrides for experiments: SLMM_METIS_OPT="index:value,index:value" (indices of METIS' options array)
  std::vector<std::pair<int, int64_t>> opt_metis;
  if (const char* env = getenv("SLMM_METIS_OPT")) {
    std::string e(env);
    size_t pos = 0;
    while (pos < e.size()) {
      size_t c = e.find(',', pos);
      if (c == std::string::npos) c = e.size();
      const std::string item = e.substr(pos, c - pos);
      const size_t colon = item.find(':');
      if (colon != std::string::npos) opt_metis.push_back({atoi(item.substr(0, colon).c_str()), atoll(item.substr(colon + 1).c_str())});
      pos = c + 1;
    }
  }
  const bool timing = getenv("SLMM_SYM_TIMING") != nullptr;
  double tp = now_s(), t_graph = 0, t_metis = 0;
  // connected components by BFS
  std::vector<int32_t> comp(n, -1), queue(n);
  std::vector<int64_t> comp_start;
  int qh = 0, qt = 0;
  ncomp = 0;
  for (int s = 0; s < n; s++) {
    if (comp[s] >= 0) continue;
    comp_start.push_back(qt);
    comp[s] = ncomp;
    queue[qt++] = s;
    while (qh < qt) {
      int u = queue[qh++];
      for (int p = ap[u]; p < ap[u + 1]; p++) {
        int v = ai[p];
        if (comp[v] < 0) { comp[v] = ncomp; queue[qt++] = v; }
      }
    }
    ncomp++;
  }
  comp_start.push_back(qt);
  if (timing) { fprintf(stderr, "[ordering] %-28s %.3f s\n", "components (BFS)", now_s() - tp); tp = now_s(); }
  // order components by size (small first) so the big fronts end up last; stable within size
  std::vector<int> corder(ncomp);
  std::iota(corder.begin(), corder.end(), 0);
  std::stable_sort(corder.begin(), corder.end(), [&](int a, int b) {
    return comp_start[a + 1] - comp_start[a] < comp_start[b + 1] - comp_start[b];
  });
  std::vector<int32_t> local(n, -1);
  int out = 0;
  std::vector<int64_t> xadj, adjncy, mperm, miperm;
  std::vector<std::vector<int>> sadj;
  std::vector<int> sorder;
  for (int ci : corder) {
    const int64_t b = comp_start[ci], e = comp_start[ci + 1];
    const int m = (int)(e - b);
    std::sort(queue.begin() + b, queue.begin() + e);      // keep original relative order inside a component
    if (m <= 3) {
      for (int64_t k = b; k < e; k++) perm[out++] = queue[k];
      continue;
    }
    for (int k = 0; k < m; k++) local[queue[b + k]] = k;
    if (m <= 256 || method == ORD_MINDEG) {
      if (m > 4096) method = ORD_METIS;      // exact MD is quadratic; large components fall through to METIS
    }
    if (m <= 256) {
      sadj.assign(m, std::vector<int>());
      for (int k = 0; k < m; k++) {
        int u = queue[b + k];
        for (int p = ap[u]; p < ap[u + 1]; p++)
          if (ai[p] != u) sadj[k].push_back(local[ai[p]]);
      }
      small_mindeg(m, sadj, sorder);
      for (int k = 0; k < m; k++) perm[out++] = queue[b + sorder[k]];
      continue;
    }
    // METIS nested dissection on the component's graph (64-bit idx_t build shipped with the CUDA toolkit)
    double tg = now_s();
    xadj.assign(m + 1, 0);
    {
      // adjacency of the component without the diagonal, 64-bit: counts, prefix sums, then the lists filled by host
      // threads over vertex ranges of equal entry counts (two passes over 6e7 entries at the 250K config)
      const int T = m < (1 << 16) ? 1 : host_threads();
      auto run = [&](auto fn) {
        std::vector<std::thread> pool;
        for (int t = 1; t < T; t++) pool.emplace_back(fn, t);
        fn(0);
        for (auto& th : pool) th.join();
      };
      run([&](int t) {
        const int k0 = (int)((int64_t)m * t / T), k1 = (int)((int64_t)m * (t + 1) / T);
        for (int k = k0; k < k1; k++) {
          const int u = queue[b + k];
          int64_t c = 0;
          for (int p = ap[u]; p < ap[u + 1]; p++)
            if (ai[p] != u) c++;
          xadj[k + 1] = c;
        }
      });
      for (int k = 0; k < m; k++) xadj[k + 1] += xadj[k];
      adjncy.resize(xadj[m]);
      run([&](int t) {
        const int k0 = (int)(std::lower_bound(xadj.begin(), xadj.end(), xadj[m] * t / T) - xadj.begin());
        const int k1 = t + 1 == T ? m : (int)(std::lower_bound(xadj.begin(), xadj.end(), xadj[m] * (t + 1) / T) - xadj.begin());
        for (int k = std::min(k0, m); k < std::min(k1, m); k++) {
          const int u = queue[b + k];
          int64_t c = xadj[k];
          for (int p = ap[u]; p < ap[u + 1]; p++)
            if (ai[p] != u) adjncy[c++] = local[ai[p]];
        }
      });
    }
    mperm.resize(m);
    miperm.resize(m);
    int64_t nv = m;
    int64_t options[40];
    METIS_SetDefaultOptions(options);
    // Tighter separator balance (UFACTOR 10 instead of METIS' 200 for node ND) and the best of 3 separators per
    // bisection (NSEPS): on the pedigree patterns of the BASELINE configs this cuts the factorization flops by
    // 10-25 % (250K: 5.12e12 -> 4.42e12, 100K: 1.83e11 -> 1.38e11) for ~20 % more ordering time, which is paid
    // once per pattern.  Indices are those of METIS 5.1's options array.
    options[16] = 10;   // METIS_OPTION_UFACTOR
    // ORD_METIS_FAST keeps one separator per bisection: at the 250K config the analysis is 30 % shorter and the
    // factorization costs 2.1 % more flops (4.51e12) - the better trade for a single fit (~36 evaluations), the
    // worse one for the evaluation time itself or for repeated fits on one pattern.
    options[15] = fast_nd ? 1 : 3;    // METIS_OPTION_NSEPS
    for (const auto& kv : opt_metis) options[kv.first] = kv.second;
    t_graph += now_s() - tg;
    tg = now_s();
    int rc = METIS_NodeND(&nv, xadj.data(), adjncy.data(), nullptr, options, mperm.data(), miperm.data());
    t_metis += now_s() - tg;
    if (rc != 1) throw std::runtime_error("METIS_NodeND failed");
    // METIS: A' = A(perm, perm); perm[new] = old
    for (int k = 0; k < m; k++) perm[out++] = queue[b + mperm[k]];
  }
  if (timing) fprintf(stderr, "[ordering] graph build %.3f s, METIS_NodeND %.3f s, rest %.3f s\n", t_graph, t_metis, now_s() - tp - t_graph - t_metis);
  if (out != n) throw std::runtime_error("ordering lost vertices");
}

// permuted full symmetric pattern, sorted rows, without the diagonal.
// No sorting: the pattern is symmetric, so the permuted matrix equals its transpose, and a transpose built by walking
// the new rows i = 0..n-1 in order and appending i to the list of every new column j it touches comes out with
// sorted lists (one counting pass + one scatter pass over the entries instead of 6e7 entries of std::sort).
void permute_pattern(int n, const int32_t* ap, const int32_t* ai, const std::vector<int32_t>& perm,
                     const std::vector<int32_t>& iperm, std::vector<int64_t>& bp, std::vector<int32_t>& bi) {
  bp.assign(n + 1, 0);
  for (int i = 0; i < n; i++) {            // by symmetry the list of new column i is as long as new row i
    const int o = perm[i];
    int64_t c = 0;
    for (int p = ap[o]; p < ap[o + 1]; p++)
      if (ai[p] != o) c++;
    bp[i + 1] = bp[i] + c;
  }
  bi.resize(bp[n]);
  const int T = (bp[n] < ((int64_t)1 << 20) && !getenv("SLMM_HOST_THREADS")) ? 1 : host_threads();
  if (T < 6) {
    // few cores: one counting pass + one scatter pass, rows taken in new order so every list comes out sorted.  The
    // scatter writes one cache line per entry (memory bound: it does not speed up with threads).
    std::vector<int64_t> fill(bp.begin(), bp.end() - 1);
    for (int i = 0; i < n; i++) {
      const int o = perm[i];
      for (int p = ap[o]; p < ap[o + 1]; p++) {
        if (ai[p] == o) continue;
        const int j = iperm[ai[p]];
        if (fill[j] >= bp[j + 1]) throw std::runtime_error("the pattern is not structurally symmetric");
        bi[fill[j]++] = i;
      }
    }
    for (int j = 0; j < n; j++)
      if (fill[j] != bp[j + 1]) throw std::runtime_error("the pattern is not structurally symmetric");
    return;
  }
  // many cores: by symmetry the list of new column j is the neighbour list of the original vertex perm[j],
  // relabelled and sorted; columns are independent, threads take column ranges of equal entry counts (sequential
  // reads, cache-resident sorts, sequential writes: 3.9 s on one core, 0.8 s on 8, against 1.05 s for the scatter).
  // Structural symmetry, which this construction relies on, is verified with order-independent hash sums of the
  // entries above / below the diagonal under the key (min, max): equal totals <=> every (i, j) has its (j, i).
  std::vector<int> bad(T, 0);
  std::vector<uint64_t> hup(T, 0), hlo(T, 0);
  auto mix = [](uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
  };
  auto work = [&](int t) {
    const int64_t lo = bp[n] * t / T, hi = bp[n] * (t + 1) / T;
    const int j0 = (int)(std::lower_bound(bp.begin(), bp.end(), lo) - bp.begin());
    const int j1 = t + 1 == T ? n : (int)(std::lower_bound(bp.begin(), bp.end(), hi) - bp.begin());
    uint64_t up = 0, lw = 0;
    for (int j = j0; j < j1; j++) {
      const int o = perm[j];
      int32_t* dst = bi.data() + bp[j];
      int64_t c = 0;
      for (int p = ap[o]; p < ap[o + 1]; p++)
        if (ai[p] != o) dst[c++] = iperm[ai[p]];
      std::sort(dst, dst + c);
      for (int64_t q = 0; q < c; q++) {
        if (q > 0 && dst[q] == dst[q - 1]) bad[t] = 1;         // duplicate entries in a row
        const uint32_t i = (uint32_t)dst[q];
        const uint64_t h = mix(((uint64_t)std::min<uint32_t>(i, (uint32_t)j) << 32) | std::max<uint32_t>(i, (uint32_t)j));
        if ((int)i > j) up += h; else lw += h;
      }
    }
    hup[t] = up;
    hlo[t] = lw;
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < T; t++) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  uint64_t su = 0, sl = 0;
  for (int t = 0; t < T; t++) {
    if (bad[t]) throw std::runtime_error("duplicate entries in the pattern");
    su += hup[t];
    sl += hlo[t];
  }
  if (su != sl) throw std::runtime_error("the pattern is not structurally symmetric");
}

// elimination tree of the permuted matrix straight from the original pattern (Liu's algorithm needs the entries of a
// column in no particular order, so nothing has to be permuted or sorted for it)
void etree_permuted(int n, const int32_t* ap, const int32_t* ai, const std::vector<int32_t>& perm,
                    const std::vector<int32_t>& iperm, std::vector<int32_t>& parent) {
  parent.assign(n, -1);
  std::vector<int32_t> anc(n, -1);
  for (int j = 0; j < n; j++) {
    const int o = perm[j];
    for (int p = ap[o]; p < ap[o + 1]; p++) {
      int i = iperm[ai[p]];
      while (i != -1 && i < j) {
        int nx = anc[i];
        anc[i] = j;
        if (nx == -1) parent[i] = j;
        i = nx;
      }
    }
  }
}

void etree(int n, const std::vector<int64_t>& bp, const std::vector<int32_t>& bi, std::vector<int32_t>& parent) {
  parent.assign(n, -1);
  std::vector<int32_t> anc(n, -1);
  for (int j = 0; j < n; j++) {
    for (int64_t p = bp[j]; p < bp[j + 1]; p++) {
      int i = bi[p];
      if (i >= j) break;                     // sorted: only the entries above the diagonal
      while (i != -1 && i < j) {
        int nx = anc[i];
        anc[i] = j;
        if (nx == -1) parent[i] = j;
        i = nx;
      }
    }
  }
}

void postorder(int n, const std::vector<int32_t>& parent, std::vector<int32_t>& post) {
  std::vector<int32_t> head(n, -1), next(n, -1), stack;
  for (int j = n - 1; j >= 0; j--)
    if (parent[j] != -1) { next[j] = head[parent[j]]; head[parent[j]] = j; }
  post.clear();
  post.reserve(n);
  for (int r = 0; r < n; r++) {
    if (parent[r] != -1) continue;
    stack.push_back(r);
    while (!stack.empty()) {
      int v = stack.back();
      int c = head[v];
      if (c == -1) { post.push_back(v); stack.pop_back(); }
      else { head[v] = next[c]; stack.push_back(c); }
    }
  }
}

// Gilbert-Ng-Peyton column counts via row-subtree leaves (skeleton matrix) with path-compressed LCA.
void column_counts(int n, const std::vector<int64_t>& bp, const std::vector<int32_t>& bi,
                   const std::vector<int32_t>& parent, std::vector<int32_t>& cc) {
  // matrix is already postordered: post[k] = k
  std::vector<int32_t> first(n, -1), maxfirst(n, -1), prevleaf(n, -1), anc(n);
  std::vector<int64_t> delta(n, 0);
  std::iota(anc.begin(), anc.end(), 0);
  for (int k = 0; k < n; k++) {
    int j = k;
    delta[j] = (first[j] == -1) ? 1 : 0;
    for (; j != -1 && first[j] == -1; j = parent[j]) first[j] = k;
  }
  for (int j = 0; j < n; j++) {
    if (parent[j] != -1) delta[parent[j]]--;
    for (int64_t p = bp[j]; p < bp[j + 1]; p++) {
      int i = bi[p];
      if (i <= j || first[j] <= maxfirst[i]) continue;
      maxfirst[i] = first[j];
      int jprev = prevleaf[i];
      prevleaf[i] = j;
      if (jprev == -1) { delta[j]++; continue; }
      int q = jprev;
      while (q != anc[q]) q = anc[q];
      for (int s = jprev; s != q;) { int sp = anc[s]; anc[s] = q; s = sp; }
      delta[j]++;
      delta[q]--;
    }
    if (parent[j] != -1) anc[j] = parent[j];
  }
  for (int j = 0; j < n; j++)
    if (parent[j] != -1) delta[parent[j]] += delta[j];
  cc.resize(n);
  for (int j = 0; j < n; j++) cc[j] = (int32_t)delta[j];
}

}  // namespace

void analyze(int n, const int32_t* ap, const int32_t* ai, const int32_t* user_perm, const SymbolicOptions& opt_in,
             Symbolic& S) {
  // relaxed-amalgamation overrides for experiments: SLMM_RELAX="n0,n1,n2,z0,z1,z2"
  SymbolicOptions opt_local = opt_in;
  if (const char* env = getenv("SLMM_RELAX")) {
    double v[6];
    if (sscanf(env, "%lf,%lf,%lf,%lf,%lf,%lf", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5]) == 6) {
      for (int q = 0; q < 3; q++) { opt_local.nrelax[q] = (int)v[q]; opt_local.zrelax[q] = v[3 + q]; }
    }
  }
  const SymbolicOptions& opt = opt_local;

  double t0 = now_s();
  S = Symbolic();
  S.n = n;
  std::vector<int32_t> perm0;
  if (opt.ordering == ORD_GIVEN) {
    if (!user_perm) throw std::runtime_error("ORD_GIVEN needs a permutation");
    perm0.assign(user_perm, user_perm + n);
    std::vector<char> seen(n, 0);
    for (int i = 0; i < n; i++) {
      if (perm0[i] < 0 || perm0[i] >= n || seen[perm0[i]]) throw std::runtime_error("invalid permutation");
      seen[perm0[i]] = 1;
    }
  } else {
    compute_ordering(n, ap, ai, opt.ordering, perm0, S.ncomponents);
  }
  S.t_order = now_s() - t0;
  t0 = now_s();
  const bool timing = getenv("SLMM_SYM_TIMING") != nullptr;
  double tp = now_s();
  auto lap = [&](const char* what) { if (timing) { fprintf(stderr, "[symbolic] %-28s %.3f s\n", what, now_s() - tp); tp = now_s(); } };

  std::vector<int32_t> iperm0(n);
  for (int i = 0; i < n; i++) iperm0[perm0[i]] = i;
  std::vector<int64_t> bp;
  std::vector<int32_t> bi, par0, post;
  const bool keep_order = (opt.ordering == ORD_GIVEN || opt.ordering == ORD_NATURAL);
  etree_permuted(n, ap, ai, perm0, iperm0, par0);
  postorder(n, par0, post);
  lap("etree + postorder");
  bool is_post = true;
  for (int k = 0; k < n; k++) if (post[k] != k) { is_post = false; break; }
  if (is_post || keep_order) {
    permute_pattern(n, ap, ai, perm0, iperm0, bp, bi);
    lap("permute_pattern");
    // A user-supplied / natural order is kept verbatim (parity mode: L is unique given P).  It is only
    // usable directly when it is already a postorder of its own etree; otherwise supernodes degrade to
    // those contiguous column runs that are chains, which is still correct.
    S.perm = perm0;
    S.iperm = iperm0;
    S.parent = par0;
  } else {
    S.perm.resize(n);
    for (int k = 0; k < n; k++) S.perm[k] = perm0[post[k]];
    S.iperm.resize(n);
    for (int i = 0; i < n; i++) S.iperm[S.perm[i]] = i;
    permute_pattern(n, ap, ai, S.perm, S.iperm, bp, bi);
    lap("permute_pattern #2");
    // the tree of the postordered matrix is the first tree relabelled (a postorder is an equivalent reordering)
    std::vector<int32_t> ipost(n);
    for (int k = 0; k < n; k++) ipost[post[k]] = k;
    S.parent.assign(n, -1);
    for (int k = 0; k < n; k++) S.parent[k] = par0[post[k]] == -1 ? -1 : ipost[par0[post[k]]];
    is_post = true;
    lap("etree #2 (relabelled)");
  }
  if (is_post) {
    column_counts(n, bp, bi, S.parent, S.colcount);
  } else {
    // column counts need a postordered tree: compute them in postorder labels and map back
    std::vector<int32_t> ipost(n);
    for (int k = 0; k < n; k++) ipost[post[k]] = k;
    std::vector<int32_t> pperm(n), piperm(n), ppar, pcc;
    for (int k = 0; k < n; k++) pperm[k] = S.perm[post[k]];
    for (int i = 0; i < n; i++) piperm[pperm[i]] = i;
    std::vector<int64_t> pbp;
    std::vector<int32_t> pbi;
    permute_pattern(n, ap, ai, pperm, piperm, pbp, pbi);
    etree(n, pbp, pbi, ppar);
    column_counts(n, pbp, pbi, ppar, pcc);
    S.colcount.resize(n);
    for (int j = 0; j < n; j++) S.colcount[j] = pcc[ipost[j]];
  }
  lap("column counts");
  const std::vector<int32_t>& parent = S.parent;
  const std::vector<int32_t>& cc = S.colcount;
  S.nnzL = 0;
  S.flops = 0;
  for (int j = 0; j < n; j++) { S.nnzL += cc[j]; S.flops += (double)cc[j] * cc[j]; }

  // ---- fundamental supernodes: column j+1 joins j when it is j's parent, j is its only child and the
  //      structures nest (count drops by exactly one).
  std::vector<int32_t> nchild(n, 0);
  for (int j = 0; j < n; j++) if (parent[j] != -1) nchild[parent[j]]++;
  std::vector<int32_t> fs_first;          // first column of each fundamental supernode
  for (int j = 0; j < n; j++) {
    bool join = j > 0 && parent[j - 1] == j && nchild[j] == 1 && cc[j - 1] == cc[j] + 1;
    if (!join) fs_first.push_back(j);
  }
  int nf = (int)fs_first.size();
  fs_first.push_back(n);
  std::vector<int32_t> col2f(n);
  for (int s = 0; s < nf; s++) for (int j = fs_first[s]; j < fs_first[s + 1]; j++) col2f[j] = s;
  std::vector<int32_t> fpar(nf, -1);
  for (int s = 0; s < nf; s++) {
    int last = fs_first[s + 1] - 1;
    fpar[s] = parent[last] == -1 ? -1 : col2f[parent[last]];
  }
  // ---- relaxed amalgamation: a supernode may absorb the child that sits immediately before it.
  //      merged[s] = representative (the absorbing ancestor); tracked with first-column / zero counts.
  std::vector<int32_t> first(nf), ncols(nf), lnz(nf);      // lnz = rows of the first column (colcount)
  std::vector<double> zeros(nf, 0.0);
  std::vector<char> alive(nf, 1);
  for (int s = 0; s < nf; s++) { first[s] = fs_first[s]; ncols[s] = fs_first[s + 1] - fs_first[s]; lnz[s] = cc[fs_first[s]]; }
  std::vector<int32_t> prev_alive(nf);    // supernode whose columns end right before first[s]
  // map "column c" -> alive supernode ending at c-1, maintained lazily through col2f + absorbed-into links
  std::vector<int32_t> into(nf);
  std::iota(into.begin(), into.end(), 0);
  auto find = [&](int s) { while (into[s] != s) { into[s] = into[into[s]]; s = into[s]; } return s; };
  const int maxcols = opt.max_super_cols > 0 ? opt.max_super_cols : (1 << 30);
  for (int p = 0; p < nf; p++) {
    if (!alive[p]) continue;
    while (first[p] > 0) {
      int c = find(col2f[first[p] - 1]);
      if (c == p) break;
      int cp = fpar[c] == -1 ? -1 : find(fpar[c]);
      if (cp != p) break;
      int ns0 = ncols[c], ns1 = ncols[p];
      int l0 = lnz[c], l1 = lnz[p];
      double newz = (double)ns0 * (ns0 + l1 - l0);
      double totz = zeros[c] + zeros[p] + newz;
      int ns = ns0 + ns1;
      double tot = (double)ns * (ns + 1) / 2 + (double)ns * (l1 - ns1);
      double z = tot > 0 ? totz / tot : 0;
      bool merge = (ns <= opt.nrelax[0]) || (ns <= opt.nrelax[1] && z < opt.zrelax[0]) ||
                   (ns <= opt.nrelax[2] && z < opt.zrelax[1]) || (z < opt.zrelax[2]) || newz == 0;
      if (!merge || ns > maxcols) break;
      alive[c] = 0;
      into[c] = p;
      first[p] = first[c];
      ncols[p] = ns;
      lnz[p] = ns0 + l1;
      zeros[p] = totz;
    }
  }
  // ---- final partition
  S.sn_first.clear();
  std::vector<int32_t> order_alive;
  for (int s = 0; s < nf; s++) if (alive[s]) order_alive.push_back(s);
  std::sort(order_alive.begin(), order_alive.end(), [&](int a, int b) { return first[a] < first[b]; });
  S.nsuper = (int)order_alive.size();
  std::vector<int32_t> newid(nf, -1);
  for (int k = 0; k < S.nsuper; k++) { newid[order_alive[k]] = k; S.sn_first.push_back(first[order_alive[k]]); }
  S.sn_first.push_back(n);
  S.col2sn.resize(n);
  for (int s = 0; s < S.nsuper; s++) for (int j = S.sn_first[s]; j < S.sn_first[s + 1]; j++) S.col2sn[j] = s;
  S.sn_parent.assign(S.nsuper, -1);
  for (int s = 0; s < S.nsuper; s++) {
    int last = S.sn_first[s + 1] - 1;
    S.sn_parent[s] = parent[last] == -1 ? -1 : S.col2sn[parent[last]];
  }
  // children lists
  S.child_ptr.assign(S.nsuper + 1, 0);
  for (int s = 0; s < S.nsuper; s++) if (S.sn_parent[s] >= 0) S.child_ptr[S.sn_parent[s] + 1]++;
  for (int s = 0; s < S.nsuper; s++) S.child_ptr[s + 1] += S.child_ptr[s];
  S.child_idx.resize(S.child_ptr[S.nsuper]);
  {
    std::vector<int32_t> fill(S.child_ptr.begin(), S.child_ptr.end() - 1);
    for (int s = 0; s < S.nsuper; s++) if (S.sn_parent[s] >= 0) S.child_idx[fill[S.sn_parent[s]]++] = s;
  }
  lap("supernodes + amalgamation");
  // ---- supernodal row structures (children before parents: supernode ids ascend with columns)
  S.sn_rowptr.assign(S.nsuper + 1, 0);
  S.sn_nrow.assign(S.nsuper, 0);
  std::vector<std::vector<int32_t>> below(S.nsuper);
  std::vector<int32_t> mark(n, -1);
  for (int s = 0; s < S.nsuper; s++) {
    const int f = S.sn_first[s], l = S.sn_first[s + 1] - 1;
    std::vector<int32_t>& b = below[s];
    for (int j = f; j <= l; j++)
      for (int64_t p = bp[j + 1] - 1; p >= bp[j]; p--) {
        int i = bi[p];
        if (i <= l) break;
        if (mark[i] != s) { mark[i] = s; b.push_back(i); }
      }
    for (int q = S.child_ptr[s]; q < S.child_ptr[s + 1]; q++) {
      std::vector<int32_t>& cb = below[S.child_idx[q]];
      for (int i : cb)
        if (i > l && mark[i] != s) { mark[i] = s; b.push_back(i); }
    }
    std::sort(b.begin(), b.end());
    S.sn_nrow[s] = (l - f + 1) + (int)b.size();
    S.sn_rowptr[s + 1] = S.sn_rowptr[s] + S.sn_nrow[s];
  }
  S.rows.resize(S.sn_rowptr[S.nsuper]);
  S.rel.assign(S.rows.size(), -1);
  S.sn_lptr.assign(S.nsuper + 1, 0);
  for (int s = 0; s < S.nsuper; s++) {
    const int f = S.sn_first[s], ns = S.sn_first[s + 1] - f;
    int64_t q = S.sn_rowptr[s];
    for (int j = 0; j < ns; j++) S.rows[q++] = f + j;
    for (int i : below[s]) S.rows[q++] = i;
    S.sn_lptr[s + 1] = S.sn_lptr[s] + panel_ld(S.sn_nrow[s]) * ns;
    S.max_front_rows = std::max(S.max_front_rows, S.sn_nrow[s]);
    S.max_super_cols = std::max(S.max_super_cols, ns);
  }
  S.lsize = S.sn_lptr[S.nsuper];
  lap("row structures");
  // relative indices into the parent's row list (merge of two sorted lists)
  for (int s = 0; s < S.nsuper; s++) {
    int p = S.sn_parent[s];
    if (p < 0) continue;
    const int ns = S.sn_first[s + 1] - S.sn_first[s];
    int64_t a = S.sn_rowptr[s] + ns, ae = S.sn_rowptr[s + 1];
    int64_t b = S.sn_rowptr[p], be = S.sn_rowptr[p + 1];
    for (; a < ae; a++) {
      while (b < be && S.rows[b] < S.rows[a]) b++;
      if (b == be || S.rows[b] != S.rows[a]) throw std::runtime_error("row structure is not nested in the parent");
      S.rel[a] = (int32_t)(b - S.sn_rowptr[p]);
    }
  }
  lap("relative indices");
  // ---- depth levels
  S.sn_depth.assign(S.nsuper, 0);
  int maxd = 0;
  for (int s = S.nsuper - 1; s >= 0; s--) {
    S.sn_depth[s] = S.sn_parent[s] < 0 ? 0 : S.sn_depth[S.sn_parent[s]] + 1;
    maxd = std::max(maxd, S.sn_depth[s]);
  }
  S.nlevels = S.nsuper ? maxd + 1 : 0;
  S.level_ptr.assign(S.nlevels + 1, 0);
  for (int s = 0; s < S.nsuper; s++) S.level_ptr[S.sn_depth[s] + 1]++;
  for (int d = 0; d < S.nlevels; d++) S.level_ptr[d + 1] += S.level_ptr[d];
  S.level_sn.resize(S.nsuper);
  {
    std::vector<int32_t> fill(S.level_ptr.begin(), S.level_ptr.end() - 1);
    for (int s = 0; s < S.nsuper; s++) S.level_sn[fill[S.sn_depth[s]]++] = s;
  }
  S.t_symbolic = now_s() - t0;
}

void entry_map(const Symbolic& S, const int32_t* ap, const int32_t* ai, int64_t* target, int tri) {
  // tri == 0: both triangles are stored with identical values (verified by the caller): the copy that lands on or
  //           below the diagonal AFTER the permutation carries the value, its mirror image is skipped.
  // tri != 0: CHOLMOD semantics - only one stored triangle defines the matrix (sksparse reads the lower triangle of
  //           the CSC matrix it is given): tri > 0 keeps the entries with index >= row of the given arrays (the lower
  //           triangle when the arrays are CSC), tri < 0 those with index <= row (the lower triangle of a CSR
  //           matrix); each kept entry is placed at (max, min) of its permuted coordinates.
  const int n = S.n;
  for (int r = 0; r < n; r++) {
    const int pr = S.iperm[r];
    for (int p = ap[r]; p < ap[r + 1]; p++) {
      const int c = ai[p];
      const int pc = S.iperm[c];
      int ir = pr, ic = pc;
      if (tri == 0) {
        if (ic > ir) { target[p] = -1; continue; }    // mirrored copy carries the value
      } else {
        if ((tri > 0 && c < r) || (tri < 0 && c > r)) { target[p] = -1; continue; }
        if (ic > ir) { ir = pc; ic = pr; }
      }
      // entry (row ir, col ic) with ir >= ic lives in the panel of ic's supernode
      const int s = S.col2sn[ic];
      const int f = S.sn_first[s];
      const int32_t* rb = S.rows.data() + S.sn_rowptr[s];
      const int32_t* re = S.rows.data() + S.sn_rowptr[s + 1];
      const int32_t* it = std::lower_bound(rb, re, ir);
      if (it == re || *it != ir) throw std::runtime_error("matrix entry outside the analysed pattern");
      target[p] = S.sn_lptr[s] + (int64_t)(ic - f) * panel_ld(S.sn_nrow[s]) + (it - rb);
    }
  }
}

}  // namespace slmm
