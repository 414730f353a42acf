"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle and the golden vectors frozen
from the unmodified reference.  Tolerances are the north-star ones: logdet / solves rel 1e-10, HE rel 1e-9,
final variance components rel 1e-6, patterns / indexing bit-exact."""
import numpy as np
import pytest
import scipy.linalg as la
import scipy.sparse as sp

from oracle import estimation as orc
from oracle.cpu_factor import DenseFactor
from tests.util import rel_err

pytestmark = pytest.mark.gpu

CASES = ["golden_small", "golden_c1mini"]


@pytest.fixture(scope="module")
def slmm():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import scilmm_b200
    from scilmm_b200 import SparseCholesky as SCmod  # noqa: F401
    import sys
    return sys.modules["scilmm_b200.SparseCholesky"]


@pytest.fixture(scope="module")
def eng():
    from scilmm_b200 import engine
    return engine


# ------------------------------------------------------------------------------------------ dense tiles
@pytest.mark.parametrize("M,N,K,lower", [(64, 64, 64, False), (128, 128, 16, False), (200, 150, 37, False),
                                         (513, 257, 130, False), (300, 300, 64, True), (1000, 40, 64, False),
                                         (7, 5, 3, False), (140, 777, 64, False), (128, 200, 5000, False),
                                         (140, 60, 20011, False), (1024, 1024, 1024, False), (2000, 1500, 300, True),
                                         (640, 384, 96, False), (258, 130, 64, False)])
def test_dmma_gemm_tiles(eng, M, N, K, lower):
    err, ms, tf = eng.gemm_selftest(M, N, K, lower=lower, reps=1)
    assert err < 1e-11 * max(K, 1), (err, M, N, K)


# ------------------------------------------------------------------------------------------ sparse kernels
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("tag", ["k1", "k3", "k4"])
def test_he_moments_vs_oracle(case, tag, request, slmm):
    g = request.getfixturevalue(case)
    mats = g.mats(tag)
    rng = np.random.default_rng(1)
    y = rng.standard_normal(g.n)
    for mqs in (False, True):
        q, S, _ = slmm.he_moments(mats, y, MQS=mqs)
        qo, So = orc.he_moments(mats, y, MQS=mqs)
        assert rel_err(q, qo) < 1e-11
        assert rel_err(S, So) < 1e-11


def test_he_moments_row_shards_add_up(golden_c1mini, eng):
    import torch
    g = golden_c1mini
    ms = eng.MatSet(g.mats("k3"))
    y = eng.to_device(np.random.default_rng(2).standard_normal(g.n))
    full = ms.he_moments_device(y).clone()
    cut = g.n // 3
    a = ms.he_moments_device(y, 0, cut).clone()
    b = ms.he_moments_device(y, cut, g.n).clone()
    empty = ms.he_moments_device(y, cut, cut).clone()
    assert torch.all(empty == 0)
    assert rel_err((a + b).cpu().numpy(), full.cpu().numpy()) < 1e-12


@pytest.mark.parametrize("case", CASES)
def test_he_estimates_vs_golden(case, request, slmm):
    g = request.getfixturevalue(case)
    for tag in ("k1", "k3"):
        est = slmm.HE(g.mats(tag), g["cov"], g["y"].copy(), compute_stderr=False)
        assert rel_err(est, g["he_" + tag]) < 1e-9
        est = slmm.HE(g.mats(tag), g["cov"], g["y"].copy(), MQS=True, compute_stderr=False)
        assert rel_err(est, g["he_mqs_" + tag]) < 1e-9


@pytest.mark.parametrize("case", CASES)
def test_he_stderr_and_bivariate_vs_golden(case, request, slmm):
    g = request.getfixturevalue(case)
    for tag in ("k1", "k3"):
        np.random.seed(g.seed + 2)
        est, se = slmm.HE(g.mats(tag), g["cov"], g["y"].copy(), compute_stderr=True, sim_num=g.sim_num)
        assert rel_err(est, g["he_" + tag]) < 1e-9
        assert np.allclose(se, g["he_se_" + tag], rtol=1e-8, equal_nan=True)
    np.random.seed(g.seed + 3)
    mats = g.mats("k1")
    y2 = g["y2"].copy()
    est, se = slmm.HE(mats, g["cov"], g["y"].copy(), compute_stderr=True, sim_num=g.sim_num, y2=y2)
    assert rel_err(est, g["he_biv"]) < 1e-9
    assert np.allclose(se, g["he_biv_se"], rtol=1e-8, equal_nan=True)
    assert mats[0].shape == (2 * g.n, 2 * g.n)          # reference mutates mat_list in place (:211)


def test_spmm_and_coldot(golden_c1mini, eng):
    g = golden_c1mini
    mats = g.mats("k3")
    ms = eng.MatSet(mats)
    rng = np.random.default_rng(4)
    for ncols in (1, 3, 33, 101, 128, 140):
        X = rng.standard_normal((g.n, ncols))
        Xd = eng.to_device(X)
        for k in (0, 2):
            ref = mats[k].dot(X)
            assert rel_err(ms.spmm(k, Xd).cpu().numpy(), ref) < 1e-13
            assert rel_err(ms.coldot(k, Xd).cpu().numpy(), np.sum(ref * X, axis=0)) < 1e-12
    x1 = rng.standard_normal(g.n)
    assert rel_err(ms.spmm(1, eng.to_device(x1)).cpu().numpy(), mats[1].dot(x1)) < 1e-13
    # fused pass over the two matrices that share a pattern (IBD and its Hadamard square)
    assert ms.pattern_id(0) == ms.pattern_id(1) != ms.pattern_id(2)
    assert sorted(map(tuple, ms.pattern_groups(2))) == [(0, 1), (2,)]
    for ncols, store_from in ((5, 2), (140, 129), (64, None), (160, 150)):
        X = rng.standard_normal((g.n, ncols))
        Xd = eng.to_device(X)
        for ks in ([0, 1], [2], [1]):
            dots, stored = ms.coldot_multi(ks, Xd, store_from)
            for gi, k in enumerate(ks):
                ref = mats[k].dot(X)
                assert rel_err(dots[gi].cpu().numpy(), np.sum(ref * X, axis=0)) < 1e-12
                if store_from is not None:
                    assert rel_err(stored[gi].cpu().numpy(), ref[:, store_from:]) < 1e-13



def test_matset_device_validation_and_pattern_sharing(golden_c1mini, eng):
    """Unsorted / duplicated CSR input is detected on the device and canonicalised; equal patterns are shared."""
    import scipy.sparse as sp
    g = golden_c1mini
    A, E_ = g.csr("A"), g.csr("E")
    rng = np.random.default_rng(2)
    # same matrix with every row's entries shuffled (has_sorted_indices is False, values are identical)
    ip = A.indptr
    perm = np.concatenate([ip[r] + rng.permutation(ip[r + 1] - ip[r]) for r in range(g.n)])
    Au = sp.csr_matrix((A.data[perm], A.indices[perm], ip.copy()), shape=A.shape)
    # duplicates: every entry split in two adjacent entries (rows non-decreasing, not strictly increasing)
    dd = np.empty(2 * A.nnz)
    dd[0::2], dd[1::2] = 0.25 * A.data, 0.75 * A.data
    Adup = sp.csr_matrix((dd, np.repeat(A.indices, 2), (2 * A.indptr).astype(np.int32)), shape=A.shape)
    y = eng.to_device(g["y"])
    ref = eng.MatSet([A, E_]).he_moments_device(y).cpu().numpy().copy()
    ms = eng.MatSet([Au, E_])
    assert ms.pattern_id(0) == ms.pattern_id(1)            # shared after canonicalisation
    assert rel_err(ms.he_moments_device(y).cpu().numpy(), ref) < 1e-12
    ms2 = eng.MatSet([Adup, E_])
    assert ms2.pattern_id(0) == ms2.pattern_id(1)
    assert rel_err(ms2.he_moments_device(y).cpu().numpy(), ref) < 1e-12
    bad = sp.csr_matrix((np.ones(1), np.array([g.n + 5], dtype=np.int32),
                         np.concatenate([[0], np.ones(g.n, dtype=np.int32)]).astype(np.int32)), shape=A.shape)
    with pytest.raises(ValueError):
        eng.MatSet([bad])


def test_quadform_symmetric_half_and_general_pass(golden_c1mini, eng):
    """x' A x per column: symmetric matrices are traversed on/below the diagonal only, a non-symmetric matrix takes
    the general pass; row shards add up (reference :65-66)."""
    import scipy.sparse as sp
    g = golden_c1mini
    mats = g.mats("k3")
    rng = np.random.default_rng(11)
    N = sp.random(g.n, g.n, density=0.01, random_state=3, format="csr") + sp.eye(g.n).tocsr()   # not symmetric
    ms = eng.MatSet(mats + [N.tocsr()])
    assert ms.is_symmetric(0) and ms.is_symmetric(1) and ms.is_symmetric(2) and not ms.is_symmetric(3)
    allm = mats + [N.tocsr()]
    for ncols in (1, 2, 31, 97, 129, 160, 200):
        X = rng.standard_normal((g.n, ncols))
        Xd = eng.to_device(X)
        for ks in ([0, 1], [2], [1], [3]):
            dots = ms.quadform_multi(ks, Xd).cpu().numpy()
            for gi, k in enumerate(ks):
                assert rel_err(dots[gi], np.sum(allm[k].dot(X) * X, axis=0)) < 1e-12
    # fused pass: wide block + narrow block whose Gram matrix rides along
    for ncols, nb in ((128, 12), (20, 1), (100, 16), (33, 5)):
        X, XB = rng.standard_normal((g.n, ncols)), rng.standard_normal((g.n, nb))
        for ks in ([0, 1], [2]):
            dots, G = ms.quadform_gram_multi(ks, eng.to_device(X), eng.to_device(XB))
            for gi, k in enumerate(ks):
                assert rel_err(dots[gi].cpu().numpy(), np.sum(allm[k].dot(X) * X, axis=0)) < 1e-12
                assert rel_err(G[gi].cpu().numpy(), XB.T.dot(allm[k].dot(XB))) < 1e-12
    X = rng.standard_normal((g.n, 40))
    Xd = eng.to_device(X)
    cut = g.n // 3
    d = ms.quadform_multi([0, 1], Xd, 0, cut) + ms.quadform_multi([0, 1], Xd, cut, g.n) + \
        ms.quadform_multi([0, 1], Xd, cut, cut)
    assert rel_err(d.cpu().numpy()[0], np.sum(mats[0].dot(X) * X, axis=0)) < 1e-12

# ------------------------------------------------------------------------------------------ factorization
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("tag", ["k2", "k4"])
@pytest.mark.parametrize("ordering", ["natural", "nesdis"])
def test_factor_protocol_vs_golden(case, tag, ordering, request, slmm):
    g = request.getfixturevalue(case)
    V = g.csc("V_" + tag)
    chol = slmm.SparseCholesky(ordering_method=ordering)
    f = chol(V)
    assert abs(f.logdet() - g["logdet_" + tag]) < 1e-10 * abs(g["logdet_" + tag])
    ys = g["y"] / g["y"].std()
    assert rel_err(f(ys), g["Viy_" + tag]) < 1e-10
    assert rel_err(f(g["cov"]), g["ViC_" + tag]) < 1e-10
    P = f.P()
    assert P.dtype == np.int32 and sorted(P.tolist()) == list(range(g.n))
    L = f.L()
    assert sp.isspmatrix_csc(L) and sp.triu(L, 1).nnz == 0
    Vp = V.toarray()[P][:, P]
    assert np.max(np.abs((L @ L.T).toarray() - Vp)) < 1e-12 * np.max(np.abs(Vp))
    if ordering == "natural":
        assert np.array_equal(P, np.arange(g.n))
        ref = DenseFactor(V)
        assert rel_err(L.toarray(), ref.L().toarray()) < 1e-11
    # second factorization with the same pattern reuses the analysis and invalidates the old Factor
    f2 = chol(V * 2.0)
    assert abs(f2.logdet() - (g["logdet_" + tag] + g.n * np.log(2.0))) < 1e-10 * abs(g["logdet_" + tag])
    with pytest.raises(RuntimeError):
        f.logdet()
    assert len(chol._engines) == 1


def test_lmul_and_probe_vectors(golden_c1mini, slmm, eng):
    g = golden_c1mini
    V = g.csc("V_k4")
    for ordering in ("natural", "nesdis"):
        f = slmm.SparseCholesky(ordering_method=ordering)(V)
        P = f.P()
        ref = DenseFactor(V, P)
        rng = np.random.default_rng(7)
        Z = rng.standard_normal((g.n, 37))
        got = f.lmul(eng.to_device(Z)).cpu().numpy()
        want = ref.L().dot(Z)[np.argsort(P)]
        assert rel_err(got, want) < 1e-11
        np.random.seed(3)
        W = slmm.simulate_vector(f, g.n, 20, None)
        np.random.seed(3)
        Wo = orc.probe_vectors(ref, np.random.randn(g.n, 20), np.argsort(P))
        assert rel_err(W, Wo) < 1e-10


def test_not_positive_definite_raises(golden_small, slmm):
    V = golden_small.csc("V_k2").tolil()
    V[5, 5] = -3.0
    with pytest.raises(slmm.NotPositiveDefiniteError):
        slmm.SparseCholesky()(V.tocsc())


def test_wide_supernodes_and_many_rhs(slmm, eng):
    # one dense 300 x 300 block (several 64-wide diagonal blocks, outer block boundary at 256) + a sparse tail
    rng = np.random.default_rng(11)
    B = rng.standard_normal((300, 300))
    D = B @ B.T + 300 * np.eye(300)
    T = sp.random(500, 500, 0.01, random_state=3)
    T = T + T.T + 30 * sp.eye(500)
    V = sp.block_diag([D, T]).tolil()
    V[300:320, 0:40] = 0.5
    V[0:40, 300:320] = 0.5
    V = V.tocsc()
    for ordering in ("natural", "nesdis"):
        f = slmm.SparseCholesky(ordering_method=ordering)(V)
        sign, ld = np.linalg.slogdet(V.toarray())
        assert abs(f.logdet() - ld) < 1e-10 * abs(ld)
        Bm = rng.standard_normal((800, 150))
        X = f(Bm)
        assert rel_err(V @ X, Bm) < 1e-10
        assert rel_err(X, np.linalg.solve(V.toarray(), Bm)) < 1e-10


@pytest.mark.parametrize("border", [300, 700])
def test_extend_add_many_children_same_targets(slmm, eng, border):
    """Block-arrow matrix: 40 leaf supernodes whose update matrices all land on the SAME entries of one parent
    front (border >= 512 rows takes the CTA-per-item extend-add, < 512 the warp-per-item one).  Lost updates between
    children show up as a wrong factor; repeated factorizations must agree bit for bit."""
    rng = np.random.default_rng(9)
    nleaf, bs = 40, 20
    n = nleaf * bs + border
    V = np.zeros((n, n))
    for b in range(nleaf):
        G = rng.standard_normal((bs, bs))
        V[b * bs:(b + 1) * bs, b * bs:(b + 1) * bs] = G @ G.T / bs + 2 * np.eye(bs)
        C = 0.05 * rng.standard_normal((border, bs))
        V[nleaf * bs:, b * bs:(b + 1) * bs] = C
        V[b * bs:(b + 1) * bs, nleaf * bs:] = C.T
    G = rng.standard_normal((border, border))
    V[nleaf * bs:, nleaf * bs:] = G @ G.T / border + 3 * np.eye(border)
    Vs = sp.csc_matrix(V)
    chol = slmm.SparseCholesky(ordering_method="natural")
    Ls = []
    for rep in range(3):
        f = chol(Vs)
        Ls.append(f.L().toarray())
    assert np.array_equal(Ls[0], Ls[1]) and np.array_equal(Ls[0], Ls[2])
    assert rel_err(Ls[0] @ Ls[0].T, V) < 1e-12
    sign, ld = np.linalg.slogdet(V)
    assert abs(f.logdet() - ld) < 1e-10 * abs(ld)
    b = rng.standard_normal((n, 3))
    assert rel_err(f(b), np.linalg.solve(V, b)) < 1e-10


@pytest.mark.parametrize("dense_first", [False, True])
def test_lookahead_streams_dense_front(slmm, eng, dense_first):
    """A 3300-column dense front: outer blocks 0..3 (1024 wide), so the trailing updates are split between the main and the bulk
    stream (look-ahead).  dense_first=True puts 300 coupled rows BELOW the wide supernode, so its Schur complement is
    built block by block on the third stream.  Factor twice and compare bit for bit (no races, fixed summation
    order), then check logdet / solve against LAPACK."""
    rng = np.random.default_rng(5)
    nd, nt, nc = 3300, 400, 300 if dense_first else 60
    B = rng.standard_normal((nd, nd))
    D = B @ B.T / nd + 2.0 * np.eye(nd)
    T = sp.random(nt, nt, 0.02, random_state=1)
    T = (T + T.T + 20 * sp.eye(nt)).toarray()
    if dense_first:
        T[:nc, :nc] += 0.01       # make the coupled rows one clique, as the fill will
    V = np.zeros((nd + nt, nd + nt))
    C = 0.01 * rng.standard_normal((nd, nc))
    if dense_first:
        V[:nd, :nd] = D
        V[nd:, nd:] = T
        V[:nd, nd:nd + nc] = C
        V[nd:nd + nc, :nd] = C.T
    else:
        V[nt:, nt:] = D
        V[:nt, :nt] = T
        V[nt:, :nc] = C
        V[:nc, nt:] = C.T
    Vs = sp.csc_matrix(V)
    chol = slmm.SparseCholesky(ordering_method="natural")
    f = chol(Vs)
    L1 = f.L().toarray()
    f = chol(Vs)
    L2 = f.L().toarray()
    assert np.array_equal(L1, L2)
    sign, ld = np.linalg.slogdet(V)
    assert abs(f.logdet() - ld) < 1e-10 * abs(ld)
    Bm = rng.standard_normal((nd + nt, 7))
    assert rel_err(f(Bm), np.linalg.solve(V, Bm)) < 1e-10
    assert rel_err(L2 @ L2.T, V) < 1e-12
    # many right-hand sides: the solve schedules use the same look-ahead (next diagonal block on the main stream,
    # the rows beyond on the bulk stream); results must be reproducible bit for bit
    Bw = rng.standard_normal((nd + nt, 96))
    X1, X2 = f(Bw), f(Bw)
    assert np.array_equal(X1, X2)
    assert rel_err(X1, np.linalg.solve(V, Bw)) < 1e-10


# ------------------------------------------------------------------------------------------ REML
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("tag", ["k2", "k4"])
def test_reml_evaluation_vs_golden(case, tag, request, slmm):
    g = request.getfixturevalue(case)
    mats, sig = g.mats(tag), g["sig_" + tag]
    ys = g["y"] / g["y"].std()
    chol = slmm.SparseCholesky(ordering_method="natural")      # golden vectors were frozen with P = identity
    for reml in (False, True):
        np.random.seed(g.seed + 4)
        nll, grad = slmm.bolt_gradient_estimation(np.log(sig), chol, mats, g["cov"], ys, reml, g.sim_num, False)
        assert abs(nll - g["bolt_nll_%s_%d" % (tag, reml)]) < 1e-10 * abs(nll)
        assert rel_err(grad, g["bolt_grad_%s_%d" % (tag, reml)]) < 1e-8
    f = chol._session(mats, g["cov"], ys).factor_at(sig)
    H = slmm.compute_hess(mats, g["cov"], f, ys)
    assert rel_err(H, g["hess_" + tag]) < 1e-9
    assert rel_err(slmm.compute_varcomp_stderr(mats, g["cov"], f, ys, g.sim_num), g["se_" + tag]) < 1e-9


def test_reml_evaluation_shared_permutation_vs_oracle(golden_c1mini, slmm):
    # engine's own nested-dissection P, oracle forced onto the same P and the same Z stream
    g = golden_c1mini
    mats, sig = g.mats("k4"), g["sig_k4"]
    ys = g["y"] / g["y"].std()
    chol = slmm.SparseCholesky()
    np.random.seed(99)
    nll, grad = slmm.bolt_gradient_estimation(np.log(sig), chol, mats, g["cov"], ys, True, 50, False)
    P = chol._session(mats, g["cov"], ys).eng.perm()
    np.random.seed(99)
    nll_o, grad_o = orc.reml_evaluation(np.log(sig), lambda M: DenseFactor(M, P), mats, g["cov"], ys, True, 50)
    assert abs(nll - nll_o) < 1e-10 * abs(nll_o)
    assert rel_err(grad, grad_o) < 1e-8


@pytest.mark.parametrize("tag,base", [("k2", "k1"), ("k4", "k3")])
def test_full_reml_fit_vs_golden(tag, base, golden_small, slmm):
    g = golden_small
    np.random.seed(g.seed + 5)
    out = slmm.REML(slmm.SparseCholesky(ordering_method="natural"), g.mats(base), g["cov"], g["y"].copy(),
                    reml=True, sim_num=g.sim_num)
    assert rel_err(out["covariance coefficients"], g["reml_sig_" + tag]) < 1e-6
    assert rel_err(out["covariates coefficients"], g["reml_beta_" + tag]) < 1e-6
    assert rel_err(out["covariance std"], g["reml_se_" + tag]) < 1e-6
    assert set(["covariance coefficients", "covariates coefficients", "covariance std"]) <= set(out)
    assert np.isfinite(out["nll"])


def test_run_estimates_dropin(golden_c1mini, slmm):
    import pandas as pd
    g = golden_c1mini
    A = g.csr("A")
    n = g.n
    cov = pd.DataFrame({"c0": g["cov"][:, 0], "c1": g["cov"][:, 1]})
    phe = pd.Series(g["y"])
    np.random.seed(5)
    est, se = slmm.run_estimates(A, phe, cov, reml=False, ignore_indices=True)
    covm = np.hstack([g["cov"][:, :2], np.ones((n, 1))])
    covm[:, :-1] -= covm[:, :-1].mean(axis=0)
    covm[:, :-1] /= covm[:, :-1].std(axis=0)
    np.random.seed(5)
    est_o, se_o = orc.he_regression([A], covm, g["y"].copy(), compute_stderr=True)
    assert rel_err(est, est_o) < 1e-9 and rel_err(se, se_o) < 1e-7


def test_legacy_lmm_and_compute_he_vs_reference(golden_small, slmm):
    """scilmm_b200.legacy.LMM / compute_HE against outputs of the unmodified reference legacy modules
    (scilmm/Estimation/LMM.py:154, HE.py:22): shared numpy probe stream, identity permutation."""
    import os
    import scilmm_b200
    from tests.util import GOLDEN_DIR
    g = golden_small
    ref = np.load(os.path.join(GOLDEN_DIR, "case_small_legacy.npz"))
    A, E_, H = g.csr("A"), g.csr("E"), g.csr("H")
    cov_raw = g["cov"][:, :-1].copy()
    for fit in (False, True):
        coef, cc = scilmm_b200.compute_HE(g["y"].copy(), cov_raw, [A, E_, H], fit_intercept=fit)
        assert rel_err(coef, ref["he_coef_%d" % fit]) < 1e-9
        assert rel_err(cc, ref["he_covcoef_%d" % fit]) < 1e-9
    chol = slmm.SparseCholesky(ordering_method="natural", rng="numpy")
    np.random.seed(21)
    out = scilmm_b200.LMM(chol, [A, E_], cov_raw, g["y"].copy(), with_intercept=True, reml=True, sim_num=20)
    assert rel_err(out["covariance coefficients"], ref["lmm_sig"]) < 1e-6
    assert rel_err(out["covariates coefficients"], ref["lmm_beta"]) < 1e-6
    assert rel_err(out["covariance std"], ref["lmm_se"]) < 1e-5
    assert rel_err(out["covariates p-values"], ref["lmm_pvalues"]) < 1e-5


def test_fused_assembly_bit_identical(golden_c1mini, slmm, eng):
    """V assembled with one pass over two same-pattern matrices (slmm_chol_add_values2) equals, bit for bit, the
    two-pass assembly and scipy's left-to-right matrices_weighted_sum (reference :55-59) scattered into the panels."""
    g = golden_c1mini
    mats = g.mats("k4")                       # A, A o A (same pattern), H, I
    sig = g["sig_k4"]
    chol = slmm.SparseCholesky(ordering_method="natural")
    ses = chol._session(mats, g["cov"], g["y"] / g["y"].std())
    assert ses.map_ids[0] == ses.map_ids[1]
    e = ses.eng
    e.add_values2(ses.map_ids[0], ses.matset.values_ptr(0), float(sig[0]), ses.matset.values_ptr(1), float(sig[1]), True)
    for k in (2, 3):
        e.add_values(ses.map_ids[k], ses.matset.values_ptr(k), float(sig[k]), False)
    fused = e.panels().copy()
    for k in range(4):
        e.add_values(ses.map_ids[k], ses.matset.values_ptr(k), float(sig[k]), k == 0)
    assert np.array_equal(fused, e.panels())
    V = slmm.matrices_weighted_sum(mats, sig)
    tgt = eng.SymbolicView(ses.union, ordering="natural").entry_map(sp.csr_matrix(V))
    Vc = eng.canonical_csr(sp.csr_matrix(V))
    ref = np.zeros_like(fused)
    ok = tgt >= 0
    ref[tgt[ok]] = Vc.data[ok]
    assert np.array_equal(fused, ref)


def test_pedigree_scale_factor_vs_oracle_panels(slmm, eng):
    """Scale guard (the extend-add race of round 1 only showed beyond ~50K individuals): simulated pedigree of 60K,
    K = 3, nested-dissection ordering.  The factor must be bit-reproducible, agree panel by panel with the CPU
    supernodal oracle on the same analysis, and solve V x = b to rounding."""
    import bench
    from oracle.supernodal_cpu import SupernodalCPUFactor, SupernodalPlan
    from scilmm_b200 import pedigree as P
    import torch
    A, _, cov, y, info = bench.make_inputs(60000, 1e-3, 3)
    n = A.shape[0]
    mats = [A, P.epistasis(A), sp.eye(n).tocsr()]
    sig = np.array([0.3, 0.15, 0.55])
    chol = slmm.SparseCholesky(rng="device")
    ses = chol._session(mats, cov, y / y.std())
    ses.factor_at(sig)
    L1, ld1 = ses.eng.panels(), ses.eng.logdet()
    ses.factor_at(sig)
    L2, ld2 = ses.eng.panels(), ses.eng.logdet()
    assert ld1 == ld2 and np.array_equal(L1, L2)
    V = (sig[0] * mats[0] + sig[1] * mats[1] + sig[2] * mats[2]).tocsr()
    ref = SupernodalCPUFactor(V, plan=SupernodalPlan(ses.union, perm=ses.eng.perm()))
    a = ref.plan.a
    first, nrow, lptr = a['sn_first'], a['sn_nrow'], a['sn_lptr']
    worst = 0.0
    for s_ in range(len(nrow)):
        ns, ms = first[s_ + 1] - first[s_], nrow[s_]
        ld = int(a['sn_ld'][s_])
        g = L2[lptr[s_]:lptr[s_] + ld * ns].reshape((ld, ns), order='F')[:ms]
        r = ref.Lx[lptr[s_]:lptr[s_] + ld * ns].reshape((ld, ns), order='F')[:ms]
        d = np.abs(g - r)
        d[:ns][np.triu(np.ones((ns, ns), bool), 1)] = 0.0          # strict upper part of the diagonal block is unused
        worst = max(worst, float(d.max()))
    assert worst < 1e-10
    assert abs(ld2 - ref.logdet()) < 1e-10 * abs(ref.logdet())
    B = np.random.default_rng(3).standard_normal((n, 70))
    X = ses.eng.solve_(eng.to_device(B)).cpu().numpy()
    assert np.abs(V.dot(X) - B).max() < 1e-10 * np.abs(B).max()


@pytest.mark.parametrize("M,N,K,lda,ldb,offa,offb,flags,copies", [
    (512, 384, 256, 512, 384, 0, 0, 0, 1),          # plain
    (512, 384, 256, 800, 800, 0, 0, 0, 1),          # panel-like leading dimensions
    (512, 384, 256, 800, 800, 0, 0, 0, 3),          # several ops in one launch
    (512, 384, 256, 800, 800, 1, 1, 0, 1),          # operands starting on an odd element (tensor-map base rounded down)
    (513, 259, 96, 802, 802, 1, 0, 0, 2),
    (736, 236, 64, 800, 800, 64, 64, 1 | 2 | 4, 1),  # lower-masked accumulate-negate: the in-block update of a narrow front
    (700, 700, 300, 1000, 1000, 301, 301, 1 | 4, 1),  # Schur complement of a front with an odd number of columns
    (512, 384, 256, 802, 802, 0, 0, 0, 1),          # column stride not a multiple of 32 bytes
    (512, 384, 256, 1000, 1000, 0, 0, 0, 1),        # ... of 128 bytes
    (513, 259, 256, 800, 800, 0, 0, 0, 1),          # ragged edges, aligned
    (513, 259, 256, 800, 800, 1, 0, 0, 1),          # ragged edges + odd start
    (512, 384, 96, 800, 800, 0, 0, 0, 1),           # K = one pass of the stage ring
    (512, 384, 300, 800, 800, 0, 0, 0, 1),          # ragged K
    (513, 259, 96, 802, 802, 0, 0, 0, 1),
])
def test_dmma_gemm_tiles_strided_multi_op(eng, M, N, K, lda, ldb, offa, offb, flags, copies):
    """The tile GEMMs (TMA-staged and cp.async-staged) on panel-shaped operands: explicit leading dimensions, odd base
    offsets, several operations per launch, lower / accumulate / negate."""
    import ctypes as C
    import torch
    from scilmm_b200._lib import check, lib
    g = torch.Generator(device="cuda").manual_seed(1)
    Abuf = torch.randn(K * lda + offa + 8, generator=g, dtype=torch.float64, device="cuda")
    Bbuf = torch.randn(K * ldb + offb + 8, generator=g, dtype=torch.float64, device="cuda")
    ldc = M + 2
    C0 = torch.randn(copies, N, ldc, generator=g, dtype=torch.float64, device="cuda")
    Cd = C0.clone()
    check(lib().slmm_gemm_selftest_ex(M, N, K, Abuf.data_ptr() + 8 * offa, lda, Bbuf.data_ptr() + 8 * offb, ldb,
                                      Cd.data_ptr(), ldc, flags, copies))
    A = Abuf[offa:offa + K * lda].view(K, lda)[:, :M]           # A[k, i] = A(i, k)
    B = Bbuf[offb:offb + K * ldb].view(K, ldb)[:, :N]
    P = (A.t() @ B).t()                                         # [N, M]: P[j, i] = sum_k A(i,k) B(j,k)
    if flags & 4:
        P = -P
    want = C0.clone()
    upd = (want[:, :, :M] + P) if (flags & 2) else P.expand(copies, N, M)
    if flags & 1:
        mask = (torch.arange(M, device="cuda")[None, :] >= torch.arange(N, device="cuda")[:, None])
        upd = torch.where(mask, upd, want[:, :, :M])
    want[:, :, :M] = upd
    assert torch.equal(Cd[:, :, M:], C0[:, :, M:])              # nothing written outside the M rows
    err = (Cd - want).abs().max().item()
    assert err < 1e-11 * K, err
