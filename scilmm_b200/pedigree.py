"""Synthetic inputs for the SparseCholesky path: pedigree simulator, IBD (numerator) matrix,
epistasis and household matrices, phenotype.  Host-side (numpy/scipy), seeded, vectorised.

These sit *before* the hot path; they exist so the BASELINE configs
(100K ... 3M individuals) can be generated in seconds-to-minutes on the GPU box, where the reference's
per-individual Python loop (Matrices/Numerator.py:12-26) would take hours.

Semantics follow the reference:
  * simulate_pedigree  ~ Simulation/Pedigree.py:17-124 (generation sizes :17-21, census-style households
    :35-54, child->household edges with keep-rate resampling :58-67, edge removal until
    nnz(IBD) ~ n^2 * sparse_factor :72-124).  It draws from numpy's Generator(seed), not the legacy global
    stream, so pedigrees are reproducible but not stream-identical to the reference simulator.
  * numerator          ~ Matrices/Numerator.py:5-43 (Henderson/Quaas  A = T D T'  with inbreeding).
  * epistasis          ~ Matrices/Epistasis.py:1-2  (A o A).
  * quick_phenotype    ~ Simulation/Phenotype.py:24-35.
"""
import numpy as np
import scipy.sparse as sp


def generation_sizes(sample_size, gen_exp):
    """reference: Simulation/Pedigree.py:17-21."""
    count = int(np.log(sample_size * gen_exp) / np.log(gen_exp))
    sizes = np.array([int(2 * (gen_exp ** x)) for x in range(count)])
    enough = np.where(np.cumsum(sizes) > sample_size)[0][0]
    head = sizes[:enough].tolist()
    return head + [sample_size - int(np.sum(head))]


def _households(rng, gen, prev):
    """Parent sets drawn from one generation (and the one before). reference: Pedigree.py:35-54.
    Returns (p1, p2) arrays; p2 == -1 for single-parent households."""
    size = gen.size
    half = size // 2
    singles = rng.choice(gen, int(size * 0.32))
    k = int(size * 0.8 * 0.68)
    a = [singles, rng.choice(gen[:half], k)]
    b = [np.full(singles.size, -1, dtype=np.int64), rng.choice(gen[half:], k)]
    if prev is not None:
        phalf = prev.size // 2
        m = int(size * 0.2 * 0.68 * 0.5)
        a += [rng.choice(gen[:half], m), rng.choice(prev[:phalf], m)]
        b += [rng.choice(prev[phalf:], m), rng.choice(gen[half:], m)]
    p1 = np.concatenate(a).astype(np.int64)
    p2 = np.concatenate(b).astype(np.int64)
    ok = p1 != p2
    return p1[ok], p2[ok]


def ibd_pattern_count(rel):
    """nnz of the IBD matrix implied by a child->parent matrix (common-ancestor count).
    reference: Matrices/Relationship.py:38-61 (get_ancestor_matrix / get_CAM / count_IBD_nonzero)."""
    n = rel.shape[0]
    R = sp.csr_matrix(rel, dtype=np.float32)
    anc = sp.eye(n, format='csr', dtype=np.float32)
    term = R
    while term.nnz > 0:
        anc = anc + term
        term = term @ R
        term.data[:] = 1.0
    anc.data[:] = 1.0
    return int((anc @ anc.T).nnz)


def simulate_pedigree(sample_size, sparse_factor, gen_exp=1.4, init_keep_rate=0.8, seed=0,
                      remove_frac=None, tol=0.1):
    """Simulated multi-generation pedigree.

    Returns dict(rel=csr bool (child row, parent col; parents precede children), sex, generation,
    household (int id per individual, -1 for founders), remove_frac).
    `remove_frac` (fraction of the shuffled edge list dropped) is searched by bisection so that
    nnz(IBD) is within `tol` of sample_size^2 * sparse_factor, unless given.
    """
    rng = np.random.default_rng(seed)
    sizes = generation_sizes(sample_size, gen_exp)
    bounds = np.concatenate(([0], np.cumsum(sizes)))
    gens = [np.arange(bounds[i], bounds[i + 1]) for i in range(len(sizes))]
    child, parent = [], []
    household = np.full(sample_size, -1, dtype=np.int64)
    hh_base = 0
    for gi in range(1, len(gens)):
        p1, p2 = _households(rng, gens[gi - 1], None if gi == 1 else gens[gi - 2])
        pick = rng.integers(0, p1.size, gens[gi].size)
        household[gens[gi]] = hh_base + pick
        hh_base += p1.size
        c = gens[gi]
        child += [c, c[p2[pick] >= 0]]
        parent += [p1[pick], p2[pick][p2[pick] >= 0]]
    child = np.concatenate(child)
    parent = np.concatenate(parent)
    total = child.size
    sel = rng.integers(0, total, int(total * init_keep_rate))     # resample with replacement (:67)
    child, parent = child[sel], parent[sel]
    order = rng.permutation(child.size)
    child, parent = child[order], parent[order]

    def build(frac):
        k = int(child.size * frac)
        m = sp.csr_matrix((np.ones(child.size - k, dtype=np.int8), (child[k:], parent[k:])),
                          shape=(sample_size, sample_size))
        m.sum_duplicates()
        m.data[:] = 1
        # an edge removed once is removed for all its duplicates (:114 assigns False by coordinate)
        if k:
            gone = sp.csr_matrix((np.ones(k, dtype=np.int8), (child[:k], parent[:k])),
                                 shape=(sample_size, sample_size))
            m = m - m.multiply(gone.astype(bool))
            m.eliminate_zeros()
        return m.astype(bool).tocsr()

    wanted = float(sample_size) ** 2 * sparse_factor
    if remove_frac is None:
        lo, hi = 0.0, 1.0
        remove_frac = 0.3
        for _ in range(40):
            cnt = ibd_pattern_count(build(remove_frac))
            if abs(cnt - wanted) < tol * wanted:
                break
            if cnt > wanted:
                lo = remove_frac
            else:
                hi = remove_frac
            remove_frac = 0.5 * (lo + hi)
        else:
            raise RuntimeError("could not reach the requested IBD sparsity")
    rel = build(remove_frac)
    assert sp.triu(rel).nnz == 0
    sex = np.zeros(sample_size)
    generation = np.zeros(sample_size, dtype=np.int64)
    for gi, g in enumerate(gens):
        sex[g[:g.size // 2]] = 1
        generation[g] = gi
    return dict(rel=rel, sex=sex, generation=generation, household=household, remove_frac=remove_frac)


def simulate_tree(sample_size, sparse_factor, gen_exp, init_keep_rate, return_households=False):
    """The reference's pedigree simulator, STREAM-IDENTICAL: same signature, same draws from the legacy global numpy
    stream in the same order, same edge order, same Nelder-Mead search for the removal fraction - so
    `np.random.seed(s); simulate_tree(...)` returns the relationship matrix the reference returns for that seed, bit for
    bit (tests/test_generator_cpu.py checks it against matrices frozen from the unmodified reference).

    reference: Simulation/Pedigree.py:93-124 (simulate_tree) with generation_size :17-21, households :35-54,
    combine_ind_to_households :58-67, count_with_removed_edges / find_number_of_edges_to_remove :72-90.
    What changes is the cost: the per-child Python list building becomes array code and the IBD-pattern count uses
    float32 sparse products instead of boolean ones (same count).  Returns (rel, sex, generation) like the reference;
    with return_households=True also the household id of every individual (-1 for founders)."""
    from scipy.optimize import fmin
    wanted = (sample_size ** 2) * sparse_factor
    sizes = generation_sizes(sample_size, gen_exp)
    bounds = np.concatenate(([0], np.cumsum(sizes)))
    gens = [np.arange(bounds[i], bounds[i + 1]) for i in range(len(sizes))]
    child, parent = [], []
    household = np.full(sample_size, -1, dtype=np.int64)
    hh_base = 0
    for gi in range(1, len(gens)):
        g, prev = gens[gi - 1], (None if gi == 1 else gens[gi - 2])
        size = g.size
        half = int(size / 2)
        singles = np.random.choice(g, int(size * 0.32))                         # :38
        k = int(size * 0.8 * 0.68)
        s1 = np.random.choice(g[:half], k)                                      # :40
        s2 = np.random.choice(g[half:], k)                                      # :41
        p1 = [singles, s1]
        p2 = [np.full(singles.size, -1, dtype=np.int64), s2]
        if prev is not None:
            phalf = int(prev.size / 2)
            m = int(size * 0.2 * 0.68 * 0.5)
            a1 = np.random.choice(g[:half], m)                                  # :49
            b1 = np.random.choice(prev[phalf:], m)                              # :50
            a2 = np.random.choice(prev[:phalf], m)                              # :51
            b2 = np.random.choice(g[half:], m)                                  # :52
            p1 += [a1, a2]
            p2 += [b1, b2]
        p1 = np.concatenate(p1).astype(np.int64)
        p2 = np.concatenate(p2).astype(np.int64)
        ok = p1 != p2                       # the reference's x != y filters (:42,:53); never true by construction
        p1, p2 = p1[ok], p2[ok]
        pick = np.random.choice(p1.size, gens[gi].size)                         # :62
        household[gens[gi]] = hh_base + pick
        hh_base += p1.size
        c = gens[gi]
        two = p2[pick] >= 0
        # edges in the reference's order: children in order, each with its parents in household order
        cnt = 1 + two.astype(np.int64)
        cc = np.repeat(c, cnt)
        pp = np.empty(cc.size, dtype=np.int64)
        first = np.cumsum(cnt) - cnt
        pp[first] = p1[pick]
        pp[first[two] + 1] = p2[pick][two]
        child.append(cc)
        parent.append(pp)
    edges = np.stack([np.concatenate(child), np.concatenate(parent)], axis=1)
    total = edges.shape[0]
    edges = edges[np.random.choice(total, int(total * init_keep_rate))]         # :67
    rel = sp.csr_matrix((np.ones(edges.shape[0]), (edges[:, 0], edges[:, 1])), shape=(sample_size, sample_size),
                        dtype=bool)
    assert sp.triu(rel).nnz == 0
    np.random.shuffle(edges)                                                    # :107

    def removed(frac):
        k = int(edges.shape[0] * frac)
        m = rel.copy()
        if k:
            gone = sp.csr_matrix((np.ones(k, dtype=np.int8), (edges[:k, 0], edges[:k, 1])), shape=rel.shape)
            m = (m.astype(np.int8) - m.astype(np.int8).multiply(gone.astype(bool))).tocsr()
            m.eliminate_zeros()
        return m.astype(bool).tocsr()

    class _Found(Exception):
        pass

    def objective(part):                                                        # :72-81
        diff = np.abs(ibd_pattern_count(removed(float(np.asarray(part).reshape(-1)[0]))) - wanted)
        if diff < 0.1 * wanted:
            raise _Found(float(np.asarray(part).reshape(-1)[0]))
        return diff

    res = None
    try:
        fmin(objective, 0.3, disp=False)                                        # :86 (returns without a hit -> None)
    except _Found as ex:
        res = ex.args[0]
    if res is None:
        raise Exception("Did not find a good enough tree")
    rel = removed(res)
    assert np.abs(ibd_pattern_count(rel) - wanted) < 0.1 * wanted
    sex = np.zeros(sample_size)
    gen_ind = np.zeros(sample_size)
    for i, g in enumerate(gens):
        sex[g[:int(g.size / 2)]] = 1
        gen_ind[g] = i
    if return_households:
        return rel, sex, gen_ind, household
    return rel, sex, gen_ind


def quick_simulate_phenotype(ibd_L, covariate_matrix, sigma_g, fixed_effects, add_intercept=False):
    """reference: Simulation/Phenotype.py:24-35, drawing from the legacy global numpy stream like the reference."""
    n = ibd_L.shape[0]
    sim = [ibd_L.dot(np.random.randn(n)), np.random.randn(n)]
    sim = np.array(sim).T
    sim = (sim - sim.mean(axis=0)) / sim.std(axis=0)
    y = sim.dot(np.sqrt(np.array([sigma_g, 1 - sigma_g])))
    if add_intercept:
        covariate_matrix = np.hstack((covariate_matrix, np.ones((n, 1))))
    y += covariate_matrix.dot(fixed_effects)
    return (y - y.mean()) / y.std()


def numerator(rel):
    """IBD / numerator relationship matrix A = T D T' with the inbreeding correction.

    reference: Matrices/Numerator.py:5-43 (LD + create_numerator).  `rel` must be lower triangular
    (parents precede children).  Returns (A csr, T csr, D 1-D array, F 1-D array).
    T = sum_k (R/2)^k is exact in binary floating point (dyadic path weights), so it equals the
    reference's row-recursive L bit for bit; D and F follow the same recurrences level by level.
    """
    R = sp.csr_matrix(rel, dtype=np.float64)
    n = R.shape[0]
    half = R * 0.5
    T = sp.eye(n, format='csr')
    term = half
    depth = np.zeros(n, dtype=np.int64)
    level = 0
    while term.nnz > 0:
        level += 1
        depth[np.diff(term.indptr) > 0] = level          # longest ancestor chain seen so far
        T = T + term
        term = term @ half
    T = sp.csr_matrix(T)
    T.sort_indices()
    nparents = np.diff(R.indptr).astype(np.float64)
    D = np.zeros(n)
    F = np.zeros(n)
    Tsq = T.copy()
    Tsq.data **= 2
    for lv in range(level + 1):
        idx = np.where(depth == lv)[0]
        if idx.size == 0:
            continue
        D[idx] = 1 - 0.25 * (nparents[idx] + R[idx].dot(F))
        F[idx] = Tsq[idx].dot(D) - 1
    A = (T @ sp.diags(D) @ T.T).tocsr()
    A.sort_indices()
    return A, T, D, F


def epistasis(ibd):
    """reference: Matrices/Epistasis.py:1-2."""
    return sp.csr_matrix(ibd.multiply(ibd))


def household_matrix(household):
    """0/1 'same parental household' block indicator with unit diagonal.  The reference has no builder
    for it (households exist only inside its simulator, Pedigree.py:35-67); definition recorded in
    DESIGN.md."""
    household = np.asarray(household)
    n = household.size
    ids = np.where(household >= 0)[0]
    _, grp = np.unique(household[ids], return_inverse=True)
    G = sp.csr_matrix((np.ones(ids.size), (ids, grp)), shape=(n, int(grp.max()) + 1 if ids.size else 1))
    M = (G @ G.T + sp.eye(n)).tocsr()
    M.data[:] = 1.0
    M.sort_indices()
    return M


def drop_unrelated(A, *others):
    """No-relatives filter of run_estimates. reference: SparseCholesky.py:363-370."""
    keep = np.asarray(A.sum(axis=1))[:, 0] > 1
    out = []
    for M in (A,) + others:
        if sp.issparse(M):
            S = sp.csr_matrix(M[keep][:, keep])
            S.eliminate_zeros()
            S.sort_indices()
            out.append(S)
        else:
            out.append(np.asarray(M)[keep])
    return keep, out


def quick_phenotype(T, D, covariates, sigma_g, fixed_effects, rng):
    """reference: Simulation/Phenotype.py:24-35 (quick_simulate_phenotype) with ibd_L = T sqrt(D)."""
    n = T.shape[0]
    g = T.dot(np.sqrt(D) * rng.standard_normal(n))
    sim = np.stack([g, rng.standard_normal(n)], axis=1)
    sim = (sim - sim.mean(axis=0)) / sim.std(axis=0)
    y = sim.dot(np.sqrt(np.array([sigma_g, 1 - sigma_g])))
    y = y + covariates.dot(fixed_effects)
    return (y - y.mean()) / y.std()
