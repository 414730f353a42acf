"""TEST INFRASTRUCTURE ONLY — freezes golden vectors by running the UNMODIFIED reference.

Run in the authoring container (needs /root/reference):   python oracle/make_golden.py
Writes tests/golden/*.npz.  The reference functions (scilmm/SparseCholesky.py HE, REML,
bolt_gradient_estimation, compute_hess, matrices_weighted_sum, ...) are imported as they are;
CHOLMOD is replaced by oracle.cpu_factor.DenseFactor (LAPACK) with the identity permutation, which is
the only possible choice here (sksparse is not installed) and makes L unique and reproducible.

All randomness is the legacy global numpy stream, seeded explicitly before each call, exactly the
stream the reference consumes (np.random.randn in :50 and :271).
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.cpu_factor import DenseFactor  # noqa: E402
from oracle.refload import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def sibship_matrix(rel):
    """0/1 'same parental household' indicator with unit diagonal (individuals with identical,
    non-empty parent sets share a household).  The reference has no builder for it."""
    rel = sp.csr_matrix(rel)
    n = rel.shape[0]
    keys = {}
    rows, cols = [], []
    for i in range(n):
        par = tuple(rel.indices[rel.indptr[i]:rel.indptr[i + 1]].tolist())
        if par:
            keys.setdefault(par, []).append(i)
    for members in keys.values():
        for a in members:
            for b in members:
                rows.append(a)
                cols.append(b)
    M = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, n))
    M = M + sp.eye(n)
    M.data[:] = 1.0
    M = sp.csr_matrix(M)
    M.sort_indices()
    return M


def pack_csr(prefix, M, d):
    M = sp.csr_matrix(M)
    M.sort_indices()
    d[prefix + "_data"] = M.data.astype(np.float64)
    d[prefix + "_indices"] = M.indices.astype(np.int32)
    d[prefix + "_indptr"] = M.indptr.astype(np.int32)


def build_case(SC, ped, num, phe, n_sim, sf, seed, c, sim_num):
    np.random.seed(seed)
    rel, sex, gen = ped.simulate_tree(n_sim, sf, 1.4, 0.8)
    ibd, L, D = num.simple_numerator(rel)
    ibd = sp.csr_matrix(ibd)
    rel = sp.csr_matrix(rel)
    np.random.seed(seed + 1)
    cov_raw = np.random.randn(n_sim, c)
    ibd_L = L @ sp.diags(np.sqrt(D.diagonal()))
    y_full = phe.quick_simulate_phenotype(ibd_L, cov_raw, 0.4, np.arange(1, c + 1) * 0.1)
    # no-relatives filter as in run_estimates (SparseCholesky.py:363-370)
    keep = np.asarray(ibd.sum(axis=1))[:, 0] > 1
    A = sp.csr_matrix(ibd[keep][:, keep])
    A.eliminate_zeros()
    A.sort_indices()
    hh = sibship_matrix(rel)[keep][:, keep]
    hh = sp.csr_matrix(hh)
    hh.sort_indices()
    y = y_full[keep].copy()
    cov = np.hstack([cov_raw[keep], np.ones((keep.sum(), 1))])
    cov[:, :-1] -= cov[:, :-1].mean(axis=0)
    cov[:, :-1] /= cov[:, :-1].std(axis=0)
    n = A.shape[0]
    epi = sp.csr_matrix(A.multiply(A))
    epi.sort_indices()

    g = {"n": n, "seed": seed, "sim_num": sim_num}
    pack_csr("rel", rel, g)
    pack_csr("ibd_full", ibd, g)
    g["ibd_D"] = D.diagonal()
    pack_csr("ibd_Lfac", sp.csr_matrix(L), g)
    pack_csr("A", A, g)
    pack_csr("E", epi, g)
    pack_csr("H", hh, g)
    g["y"] = y
    g["cov"] = cov
    g["keep"] = keep

    chol = lambda V: DenseFactor(V)  # identity permutation

    # ---- HE (reference :192-281)
    for tag, mats in (("k1", [A]), ("k3", [A, epi, hh])):
        g["he_%s" % tag] = SC.HE(list(mats), cov, y.copy(), compute_stderr=False)
        g["he_mqs_%s" % tag] = SC.HE(list(mats), cov, y.copy(), MQS=True, compute_stderr=False)
        np.random.seed(seed + 2)
        est, se = SC.HE(list(mats), cov, y.copy(), compute_stderr=True, sim_num=sim_num)
        g["he_se_%s" % tag] = se
    # bivariate mode (:202-211)
    y2 = np.roll(y, 7) * 0.5 + 0.5 * y
    np.random.seed(seed + 3)
    est2, se2 = SC.HE([A.copy()], cov, y.copy(), compute_stderr=True, sim_num=sim_num, y2=y2.copy())
    g["y2"] = y2
    g["he_biv"] = est2
    g["he_biv_se"] = se2

    # ---- fixed-sigma REML pieces (reference :29-117), K=2 and K=4
    ys = y / y.std()
    eye = sp.eye(n).tocsr()
    for tag, mats, sig in (("k2", [A, eye], np.array([0.4, 0.6])),
                           ("k4", [A, epi, hh, eye], np.array([0.3, 0.15, 0.1, 0.45]))):
        V = SC.matrices_weighted_sum(mats, sig)
        V.sort_indices()
        f = chol(V)
        g["V_%s_data" % tag] = V.data
        g["V_%s_indices" % tag] = V.indices.astype(np.int32)
        g["V_%s_indptr" % tag] = V.indptr.astype(np.int32)
        g["sig_%s" % tag] = sig
        g["logdet_%s" % tag] = f.logdet()
        ViC, Lc, mu, beta = SC.estimate_fixed_effects(f, ys, cov)
        Vir = f(ys - mu)
        g["ViC_%s" % tag] = ViC
        g["beta_%s" % tag] = beta
        g["Viy_%s" % tag] = f(ys)
        g["Vir_%s" % tag] = Vir
        g["nll_ml_%s" % tag] = SC.negative_log_likelihood(f, ys, Vir, mu, Lc, False)
        g["nll_reml_%s" % tag] = SC.negative_log_likelihood(f, ys, Vir, mu, Lc, True)
        for reml in (False, True):
            np.random.seed(seed + 4)
            nll, grad = SC.bolt_gradient_estimation(np.log(sig), chol, mats, cov, ys, reml, sim_num, False)
            g["bolt_nll_%s_%d" % (tag, reml)] = nll
            g["bolt_grad_%s_%d" % (tag, reml)] = grad
        g["hess_%s" % tag] = SC.compute_hess(mats, cov, f, ys)
        g["se_%s" % tag] = SC.compute_varcomp_stderr(mats, cov, f, ys, sim_num)

    # ---- full REML fits under a frozen stream, identity permutation (reference :177-189)
    for tag, mats in (("k2", [A]), ("k4", [A, epi, hh])):
        np.random.seed(seed + 5)
        out = SC.REML(chol, list(mats), cov, y.copy(), reml=True, sim_num=sim_num, verbose=False)
        g["reml_sig_%s" % tag] = out["covariance coefficients"]
        g["reml_beta_%s" % tag] = out["covariates coefficients"]
        g["reml_se_%s" % tag] = out["covariance std"]

    # ---- MINQUE (reference :284-347), 2 iterations (one factorization of H; a third makes H indefinite on these inputs) under a frozen stream, identity permutation
    import contextlib
    import io
    for tag, mats in (("k1", [A]), ("k3", [A, epi, hh])):
        np.random.seed(seed + 6)
        with contextlib.redirect_stdout(io.StringIO()):          # the reference prints every iteration (:330)
            g["minque_%s" % tag] = SC.MINQUE(chol, list(mats), cov, y.copy(), num_iter=2, sim_num=sim_num)
    return g


def main():
    SC, ped, num, phe = load_reference()
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    # case_c1 is BASELINE.json config 1 at its stated size: simulate_tree(10000, 1e-3, 1.4, 0.8) -> ~7.1K individuals
    # after the no-relatives filter, 1 IBD matrix (+ the K=4 variants), 2 covariates + intercept, sim_num 100
    for name, kw in (("case_small", dict(n_sim=400, sf=0.01, seed=11, c=2, sim_num=20)),
                     ("case_c1mini", dict(n_sim=2500, sf=0.004, seed=0, c=2, sim_num=100)),
                     ("case_c1", dict(n_sim=10000, sf=0.001, seed=0, c=2, sim_num=100))):
        if only and name not in only:
            continue
        g = build_case(SC, ped, num, phe, **kw)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **g)
        print(name, "n=%d nnz(A)=%d nnz(H)=%d" % (g["n"], g["A_data"].size, g["H_data"].size),
              "he_k1", g["he_k1"], "he_k3", g["he_k3"], "reml_k2", g["reml_sig_k2"],
              "reml_k4", g["reml_sig_k4"], "%.0f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
