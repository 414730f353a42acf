"""Pins the oracle restatement (oracle/estimation.py) against outputs of the unmodified reference
frozen in tests/golden/ by oracle/make_golden.py.  CPU only."""
import numpy as np
import pytest

from oracle import estimation as orc
from oracle.cpu_factor import DenseFactor, SuperLUFactor
from tests.util import rel_err

CASES = ["golden_small", "golden_c1mini"]


@pytest.mark.parametrize("case", CASES)
def test_he_estimates(case, request):
    g = request.getfixturevalue(case)
    for tag in ("k1", "k3"):
        est = orc.he_regression(g.mats(tag), g["cov"], g["y"].copy(), compute_stderr=False)
        assert rel_err(est, g["he_" + tag]) < 1e-12
        est = orc.he_regression(g.mats(tag), g["cov"], g["y"].copy(), MQS=True, compute_stderr=False)
        assert rel_err(est, g["he_mqs_" + tag]) < 1e-12


@pytest.mark.parametrize("case", CASES)
def test_he_stderr_stream(case, request):
    g = request.getfixturevalue(case)
    for tag in ("k1", "k3"):
        np.random.seed(g.seed + 2)
        _, se = orc.he_regression(g.mats(tag), g["cov"], g["y"].copy(), compute_stderr=True, sim_num=g.sim_num)
        assert np.allclose(se, g["he_se_" + tag], rtol=1e-10, atol=0, equal_nan=True)


@pytest.mark.parametrize("case", CASES)
def test_he_bivariate(case, request):
    g = request.getfixturevalue(case)
    np.random.seed(g.seed + 3)
    est, se = orc.he_regression(g.mats("k1"), g["cov"], g["y"].copy(), compute_stderr=True,
                                sim_num=g.sim_num, y2=g["y2"].copy())
    assert rel_err(est, g["he_biv"]) < 1e-12
    assert np.allclose(se, g["he_biv_se"], rtol=1e-10, equal_nan=True)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("tag", ["k2", "k4"])
def test_fixed_sigma_pieces(case, tag, request):
    g = request.getfixturevalue(case)
    mats, sig = g.mats(tag), g["sig_" + tag]
    ys = g["y"] / g["y"].std()
    V = orc.weighted_sum(mats, sig)
    V.sort_indices()
    Vg = g.csc("V_" + tag)
    assert np.array_equal(V.indptr, Vg.indptr) and np.array_equal(V.indices, Vg.indices)
    assert np.array_equal(V.data, Vg.data)          # bit-exact: same scipy kernels, same order
    f = DenseFactor(V)
    assert abs(f.logdet() - g["logdet_" + tag]) < 1e-12 * abs(g["logdet_" + tag])
    ViC, chol, mu, beta = orc.fixed_effects(f, ys, g["cov"])
    assert rel_err(ViC, g["ViC_" + tag]) < 1e-12
    assert rel_err(beta, g["beta_" + tag]) < 1e-12
    Vir = f(ys - mu)
    assert rel_err(Vir, g["Vir_" + tag]) < 1e-12
    assert abs(orc.nll_value(f, ys, Vir, mu, chol, False) - g["nll_ml_" + tag]) < 1e-12 * abs(g["nll_ml_" + tag])
    assert abs(orc.nll_value(f, ys, Vir, mu, chol, True) - g["nll_reml_" + tag]) < 1e-12 * abs(g["nll_reml_" + tag])
    for reml in (False, True):
        np.random.seed(g.seed + 4)
        nll, grad = orc.reml_evaluation(np.log(sig), lambda M: DenseFactor(M), mats, g["cov"], ys, reml,
                                        g.sim_num)
        assert abs(nll - g["bolt_nll_%s_%d" % (tag, reml)]) < 1e-12 * abs(nll)
        assert rel_err(grad, g["bolt_grad_%s_%d" % (tag, reml)]) < 1e-10
    H = orc.average_information(mats, g["cov"], f, ys)
    assert rel_err(H, g["hess_" + tag]) < 1e-10
    assert rel_err(orc.varcomp_stderr(mats, g["cov"], f, ys, g.sim_num), g["se_" + tag]) < 1e-10


@pytest.mark.parametrize("tag,base", [("k2", "k1"), ("k4", "k3")])
def test_full_reml_fit_small(tag, base, golden_small):
    g = golden_small
    np.random.seed(g.seed + 5)
    out = orc.reml_fit(lambda M: DenseFactor(M), g.mats(base), g["cov"], g["y"].copy(), reml=True,
                       sim_num=g.sim_num)
    assert rel_err(out["covariance coefficients"], g["reml_sig_" + tag]) < 1e-8
    assert rel_err(out["covariates coefficients"], g["reml_beta_" + tag]) < 1e-8
    assert rel_err(out["covariance std"], g["reml_se_" + tag]) < 1e-8


def test_superlu_factor_matches_dense(golden_small):
    g = golden_small
    V = g.csc("V_k4")
    rng = np.random.default_rng(3)
    perm = rng.permutation(g.n).astype(np.int32)
    fd, fs = DenseFactor(V, perm), SuperLUFactor(V, perm)
    b = rng.standard_normal((g.n, 3))
    assert rel_err(fs(b), fd(b)) < 1e-11
    assert abs(fs.logdet() - fd.logdet()) < 1e-11 * abs(fd.logdet())
    assert rel_err(fs.L().toarray(), fd.L().toarray()) < 1e-11
    assert np.array_equal(fs.P(), perm)


def test_legacy_entry_points_vs_reference(golden_small):
    """oracle restatement of scilmm/Estimation/LMM.py:154 and HE.py:22 against outputs of the unmodified reference
    (oracle/make_golden_legacy.py -> tests/golden/case_small_legacy.npz)."""
    import os
    from tests.util import GOLDEN_DIR
    g = golden_small
    ref = np.load(os.path.join(GOLDEN_DIR, "case_small_legacy.npz"))
    A, E, H = g.csr("A"), g.csr("E"), g.csr("H")
    cov_raw = g["cov"][:, :-1].copy()
    for fit in (False, True):
        coef, cc = orc.legacy_compute_he(g["y"].copy(), cov_raw, [A, E, H], fit_intercept=fit)
        assert rel_err(coef, ref["he_coef_%d" % fit]) < 1e-10
        assert rel_err(cc, ref["he_covcoef_%d" % fit]) < 1e-10
    np.random.seed(21)
    out = orc.legacy_lmm(lambda V: DenseFactor(V), [A, E], cov_raw, g["y"].copy(), True, True, 20)
    assert rel_err(out["covariance coefficients"], ref["lmm_sig"]) < 1e-6
    assert rel_err(out["covariates coefficients"], ref["lmm_beta"]) < 1e-6
    assert rel_err(out["covariance std"], ref["lmm_se"]) < 1e-5
    assert rel_err(out["covariates p-values"], ref["lmm_pvalues"]) < 1e-5


@pytest.mark.parametrize("case", CASES)
def test_minque_two_iterations(case, request):
    """oracle restatement of MINQUE (reference :284-347) against the unmodified reference, frozen stream, P = I."""
    g = request.getfixturevalue(case)
    for tag in ("k1", "k3"):
        np.random.seed(g.seed + 6)
        est = orc.minque(lambda M: DenseFactor(M), g.mats(tag), g["cov"], g["y"].copy(), num_iter=2,
                         sim_num=g.sim_num)
        assert rel_err(est, g["minque_" + tag]) < 1e-9


def test_c1_at_stated_size(golden_c1):
    """Config 1 at n_sim = 10,000 (7,108 individuals kept): HE, assembly bit-exactness and one REML evaluation of the
    restatement against the unmodified reference (the dense LAPACK factor takes ~1 s at this size)."""
    g = golden_c1
    assert g.n == 7108 and g.csr("A").nnz == 105418
    for tag in ("k1", "k3"):
        est = orc.he_regression(g.mats(tag), g["cov"], g["y"].copy(), compute_stderr=False)
        assert rel_err(est, g["he_" + tag]) < 1e-12
    mats, sig = g.mats("k2"), g["sig_k2"]
    ys = g["y"] / g["y"].std()
    V = orc.weighted_sum(mats, sig)
    V.sort_indices()
    Vg = g.csc("V_k2")
    assert np.array_equal(V.indptr, Vg.indptr) and np.array_equal(V.indices, Vg.indices)
    assert np.array_equal(V.data, Vg.data)
    np.random.seed(g.seed + 4)
    nll, grad = orc.reml_evaluation(np.log(sig), lambda M: DenseFactor(M), mats, g["cov"], ys, True, g.sim_num)
    assert abs(nll - g["bolt_nll_k2_1"]) < 1e-12 * abs(nll)
    assert rel_err(grad, g["bolt_grad_k2_1"]) < 1e-9
