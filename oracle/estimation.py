"""TEST INFRASTRUCTURE ONLY — CPU oracle for the SciLMM SparseCholesky path.

A numpy/scipy restatement of the algorithm in the reference file scilmm/SparseCholesky.py.
Nothing under scilmm_b200/ may import this module; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs use it, and only as the checker / the timed CPU arm.

Parity pinning: the reference ships no golden vectors for this path.  This
restatement is pinned against outputs of the *unmodified* reference run in the authoring container
(oracle/make_golden.py -> tests/golden/*.npz) by tests/test_oracle_golden.py.

Every function cites the reference lines it follows.  The sparse kernels used are the same scipy
routines the reference calls (csr_matvec(s), csr_plus_csr, csr_elmult_csr), so timing this module
is a faithful stand-in for timing the reference's CPU arithmetic.
"""
import numpy as np
import scipy.linalg as la
import scipy.optimize as optimize
import scipy.sparse as sp

LOG_2PI = np.log(2.0 * np.pi)


# ----------------------------------------------------------------------------- V assembly
def weighted_sum(mats, sigmas):
    """V = sum_k sigma_k * mats[k], accumulated left to right, returned CSC.

    reference: SparseCholesky.py:55-59 (matrices_weighted_sum).
    """
    acc = sigmas[0] * mats[0]
    for k in range(1, len(sigmas)):
        acc = acc + sigmas[k] * mats[k]
    return acc.tocsc()


# ----------------------------------------------------------------------------- fixed effects / nll
def fixed_effects(factor, y, C):
    """GLS fixed effects. reference: SparseCholesky.py:29-34 (estimate_fixed_effects)."""
    ViC = factor(C)
    chol_CtViC = la.cho_factor(C.T.dot(ViC))
    beta = la.cho_solve(chol_CtViC, C.T.dot(factor(y)))
    return ViC, chol_CtViC, C.dot(beta), beta


def nll_value(factor, y, Vi_r, mu, chol_CtViC, reml):
    """0.5*[(y-mu)'V^-1(y-mu) + n log 2pi + logdet V] (+ sum log diag chol(C'V^-1C) for REML).

    reference: SparseCholesky.py:37-46 (negative_log_likelihood).
    """
    n = y.size
    val = 0.5 * ((y - mu).dot(Vi_r) + n * LOG_2PI + factor.logdet())
    if reml:
        val += 0.5 * 2 * np.sum(np.log(np.diag(chol_CtViC[0])))
    return val


def probe_vectors(factor, Z, p_inv):
    """W = V^-1 (L Z)[argsort(P)] for a given normal block Z (n x s).

    reference: SparseCholesky.py:49-52 (simulate_vector); there Z = np.random.randn(n, sim_num).
    """
    return factor(factor.L().dot(Z)[p_inv])


def gradient_terms(sigmas, mats, W, Vi_r, reml, ViC, chol_CtViC):
    """d nll / d sigma_k. reference: SparseCholesky.py:62-74 (compute_gradients)."""
    g = np.zeros(len(sigmas))
    for k in range(len(sigmas)):
        trace_mc = np.mean(np.sum(mats[k].dot(W) * W, axis=0))
        quad = Vi_r.dot(mats[k].dot(Vi_r))
        g[k] = 0.5 * (trace_mc - quad)
        if reml:
            inner = ViC.T.dot(mats[k].dot(ViC))
            g[k] -= 0.5 * np.trace(la.cho_solve(chol_CtViC, inner))
    return g


def reml_evaluation(log_sigmas, cholesky_func, mats, C, y, reml, sim_num, verbose=False, take_exp=True,
                    normal_source=None, detail=None):
    """One REML objective evaluation (nll, grad wrt log sigma).

    reference: SparseCholesky.py:77-117 (bolt_gradient_estimation).  `normal_source(n, s)` supplies
    the probe block; default is the global numpy stream exactly as the reference (:50).
    """
    sigmas = np.exp(log_sigmas) if take_exp else np.asarray(log_sigmas, dtype=float)
    V = weighted_sum(mats, sigmas)
    n = V.shape[0]
    factor = cholesky_func(V)
    perm = factor.P()
    p_inv = np.argsort(perm)
    ViC, chol_CtViC, mu, beta = fixed_effects(factor, y, C)
    Vi_r = factor(y - mu)
    nll = nll_value(factor, y, Vi_r, mu, chol_CtViC, reml)
    Z = np.random.randn(n, sim_num) if normal_source is None else normal_source(n, sim_num)
    W = probe_vectors(factor, Z, p_inv)
    grad = gradient_terms(sigmas, mats, W, Vi_r, reml, ViC, chol_CtViC)
    if take_exp:
        grad = grad * sigmas
    if detail is not None:
        detail.update(V=V, ViC=ViC, mu=mu, beta=beta, Vi_r=Vi_r, W=W, logdet=factor.logdet(), perm=perm)
    return nll, grad


def fit_variance_components(cholesky_func, mats, C, y, reml=True, sim_num=100, verbose=False,
                            normal_source=None, trace=None):
    """HE start + L-BFGS-B on log sigma. reference: SparseCholesky.py:120-144 (estimate_var_comps)."""
    start = he_regression(mats[:-1], C, y, compute_stderr=False)
    start = np.concatenate((start, [1 - start.sum()]))
    if np.any(start < 0):
        start = np.ones(len(mats))
    start = start / start.sum()

    def objective(x):
        out = reml_evaluation(x, cholesky_func, mats, C, y, reml, sim_num, verbose, True, normal_source)
        if trace is not None:
            trace.append((np.array(x, copy=True), out[0], np.array(out[1], copy=True)))
        return out

    res = optimize.minimize(objective, np.log(start), jac=True, method='L-BFGS-B',
                            options={'eps': 1e-5, 'ftol': 1e-7})
    return np.exp(res.x)


def average_information(mats, C, factor, y):
    """hess[i,j] = -0.5 y' P A_i P A_j P y. reference: SparseCholesky.py:147-168 (compute_hess)."""
    K = len(mats)
    ViC = factor(C)
    chol_CtViC = la.cho_factor(C.T.dot(ViC))

    def project(z):
        Viz = factor(z)
        return Viz - ViC.dot(la.cho_solve(chol_CtViC, C.T.dot(Viz)))

    Py = project(y)
    H = np.empty((K, K))
    for j in range(K):
        PAjPy = project(mats[j].dot(Py))
        for i in range(j + 1):
            H[i, j] = -0.5 * y.dot(project(mats[i].dot(PAjPy)))
            H[j, i] = H[i, j]
    return H


def varcomp_stderr(mats, C, factor, y, sim_num):
    """reference: SparseCholesky.py:171-174 (compute_varcomp_stderr)."""
    H = average_information(mats, C, factor, y)
    return np.sqrt(np.diag(la.inv(-H)) * (1 + 1.0 / sim_num))


def reml_fit(cholesky_func, mats, C, y, reml=True, sim_num=100, verbose=False, normal_source=None):
    """reference: SparseCholesky.py:177-189 (REML).  Extra key 'nll' is not in the reference dict."""
    y = y / y.std()
    mats = list(mats) + [sp.eye(y.shape[0]).tocsr()]
    sig = fit_variance_components(cholesky_func, mats, C, y, reml, sim_num, verbose, normal_source)
    factor = cholesky_func(weighted_sum(mats, sig))
    _, _, _, beta = fixed_effects(factor, y, C)
    se = varcomp_stderr(mats, C, factor, y, sim_num)
    return {"covariance coefficients": sig, "covariates coefficients": beta, "covariance std": se}


# ----------------------------------------------------------------------------- Haseman-Elston
def he_moments(mats, y, MQS=False):
    """q[i], S[i,j] of the HE normal equations. reference: SparseCholesky.py:213-243."""
    K = len(mats)
    n = y.shape[0]
    q = np.zeros(K)
    S = np.zeros((K, K))
    for i in range(K):
        Ai = mats[i]
        if MQS:
            q[i] = y.dot(Ai.dot(y)) - y.dot(y)
        elif sp.issparse(Ai):
            q[i] = y.dot(Ai.dot(y)) - Ai.diagonal().dot(y ** 2)
        else:
            q[i] = ((Ai * y).T.dot(y)).sum() - np.diag(Ai).dot(y ** 2)
        for j in range(i + 1):
            Aj = mats[j]
            if MQS:
                S[i, j] = (Ai.multiply(Aj)).sum() - (n - 1)
            elif sp.issparse(Ai):
                S[i, j] = (Ai.multiply(Aj)).sum() - Ai.diagonal().dot(Aj.diagonal())
            else:
                S[i, j] = np.einsum('ij,ij->', Ai, Aj) - np.diag(Ai).dot(np.diag(Aj))
            S[j, i] = S[i, j]
    return q, S


def he_regression(mat_list, cov, y, MQS=False, verbose=False, sim_num=100, compute_stderr=False, y2=None,
                  normal_source=None, detail=None):
    """Haseman-Elston regression. reference: SparseCholesky.py:192-281 (HE).

    Reproduces the reference including its K>1 quirks in the sampling-variance branch
    (:252-255 stale index `i`; :262-274 rebinding of mat_i and stale mat_j).
    Like the reference, bivariate mode mutates `mat_list` and `y2` in place (:203-211).
    """
    CtC = cov.T.dot(cov)
    y = y - cov.dot(np.linalg.solve(CtC, cov.T.dot(y)))
    y /= y.std()
    if y2 is not None:
        y2 -= cov.dot(np.linalg.solve(CtC, cov.T.dot(y2)))
        y2 /= y2.std()
        y = np.concatenate((y, y2))
        for idx, m in enumerate(mat_list):
            z = sp.csr_matrix((m.shape[0], m.shape[0]))
            mat_list[idx] = sp.vstack([sp.hstack([z, m]), sp.hstack([m, z])]).tocsr()

    K = len(mat_list)
    n = y.shape[0]
    q, S = he_moments(mat_list, y, MQS)
    est = np.linalg.solve(S, q)
    if detail is not None:
        detail.update(q=q, S=S, y=y)
    if not compute_stderr:
        return est

    stale_i = K - 1                     # value of the loop variable `i` after :216-232
    stale_mat_j = mat_list[K - 1]       # value of `mat_j` after the inner loop at :224
    H = mat_list[0] * est[0]
    for m in mat_list[1:]:
        H = H + m * est[stale_i]
    H = H + sp.eye(n, format='csr') * (1.0 - est.sum())

    Vq = np.empty((K, K))
    for i in range(K):
        for j in range(i + 1):
            left = mat_list[j]          # the inner loop rebinds mat_i to mat_list[j] (:262)
            if sim_num is None:         # exact branch (:260-268): sparse x sparse products
                HAi = H.dot(mat_list[i]) - H
                HAj = HAi if j == i else H.dot(stale_mat_j) - H
                Vq[i, j] = 2 * (HAi.multiply(HAj)).sum()
                Vq[j, i] = Vq[i, j]
                continue
            Zs = np.random.randn(n, sim_num) if normal_source is None else normal_source(n, sim_num)
            t1 = stale_mat_j.dot(Zs) - Zs
            t2 = H.dot(t1)
            t3 = left.dot(t2) - t2
            t4 = H.dot(t3)
            Vq[i, j] = 2 * np.mean(np.einsum('ij,ij->j', Zs, t4))
            Vq[j, i] = Vq[i, j]
    var = np.linalg.solve(S, np.linalg.solve(S, Vq).T).T
    return est, np.sqrt(np.diag(var))


def minque(cholesky_func, mat_list, cov, y, compute_stderr=False, verbose=False, num_iter=100, sim_num=100,
           normal_source=None):
    """Iterated MINQUE. reference: SparseCholesky.py:284-347.

    First pass (H = I, :296) is the MQS moments; later passes weight by H^-1 with Monte-Carlo traces through
    probes of covariance H (:303-306).  Reproduces the stale loop index in the update of H (:335: every matrix
    after the first is weighted by minque_est[K-1]).  compute_stderr returns (est, 0.0) as the reference (:343-347).
    """
    CtC = cov.T.dot(cov)
    y = y - cov.dot(np.linalg.solve(CtC, cov.T.dot(y)))
    y /= y.std()
    K = len(mat_list)
    n = y.shape[0]
    H = None
    for it in range(num_iter):
        q = np.zeros(K)
        S = np.zeros((K, K))
        if H is not None:
            factor = cholesky_func(H)
            Z = np.random.randn(n, sim_num) if normal_source is None else normal_source(n, sim_num)
            sim_y = factor.L().dot(Z)[np.argsort(factor.P())]
            Hi_sim = factor(sim_y)
            Hi_y = factor(y)
        for i in range(K):
            if H is None:
                q[i] = y.dot(mat_list[i].dot(y)) - y.dot(y)
            else:
                q[i] = Hi_y.dot(mat_list[i].dot(Hi_y)) - y.dot(y)
                Hi_Ki_Hi_sim = factor(mat_list[i].dot(Hi_sim))
            for j in range(i + 1):
                if H is None:
                    S[i, j] = (mat_list[i].multiply(mat_list[j])).sum() - (n - 1)
                else:
                    S[i, j] = np.mean(np.einsum('ij,ij->j', Hi_sim, mat_list[j].dot(Hi_Ki_Hi_sim))) - (n - 1)
                S[j, i] = S[i, j]
        est = np.linalg.solve(S, q)
        if verbose:
            print(it + 1, est)
        stale_i = K - 1
        H = mat_list[0] * est[0]
        for m in mat_list[1:]:
            H = H + m * est[stale_i]
        H = H + sp.eye(n, format='csr') * (1.0 - est.sum())
    est = np.linalg.solve(S, q)
    if not compute_stderr:
        return est
    return est, np.sqrt(0)


# ----------------------------------------------------------------------------- legacy entry points
def legacy_lmm(cholesky_func, mats, C, y, with_intercept=True, reml=True, sim_num=100, verbose=False):
    """reference: scilmm/Estimation/LMM.py:154-171 (LMM) with compute_sigmas :111-124 (equal starting components,
    no HE start, y not standardised), compute_fixed_effects_p_value :127-131 and compute_sig_of_sig :134-151
    (the same average-information matrix as SparseCholesky.py:147-168)."""
    import scipy.stats as stats
    mats = list(mats) + [sp.eye(y.size).tocsr()]
    if with_intercept:
        C = np.hstack((np.ones((y.size, 1)), C))
    x0 = np.log(np.ones(len(mats)) / len(mats))
    res = optimize.minimize(lambda x: reml_evaluation(x, cholesky_func, mats, C, y, reml, sim_num, verbose, True),
                            x0, jac=True, method='L-BFGS-B', options={'eps': 1e-5, 'ftol': 1e-7})
    sig = np.exp(res.x)
    factor = cholesky_func(weighted_sum(mats, sig))
    _, chol_CtViC, _, beta = fixed_effects(factor, y, C)
    var_beta = la.cho_solve(chol_CtViC, np.eye(C.shape[1]))
    pvals = stats.f(1, y.shape[0] - 1).sf(beta ** 2 / np.diag(var_beta))
    se = legacy_sig_of_sig(mats, C, factor, y, sim_num)
    return {"covariance coefficients": sig, "covariates coefficients": beta, "covariance std": se,
            "covariates p-values": pvals}


def legacy_sig_of_sig(mats, C, factor, y, sim_num):
    """reference: scilmm/Estimation/LMM.py:134-151 (compute_sig_of_sig).  Unlike SparseCholesky.py:147-168 only the
    innermost vector is projected: hess[i,j] = -0.5 y' V^-1 A_i V^-1 A_j P y."""
    K = len(mats)
    Viy, ViC = factor(y), factor(C)
    Py = Viy - ViC.dot(np.linalg.inv(C.T.dot(ViC)).dot(C.T.dot(Viy)))
    F = [factor(mats[j].dot(Py)) for j in range(K)]
    H = np.empty((K, K))
    for i in range(K):
        for j in range(i, K):
            H[i, j] = H[j, i] = -0.5 * y.dot(factor(mats[i].dot(F[j])))
    return np.sqrt(np.diag(la.inv(-H)) * (1 + 1.0 / sim_num))


def legacy_compute_he(y, C, mats, fit_intercept=False):
    """reference: scilmm/Estimation/HE.py:22-40 (compute_HE) with regress_beta_out :8-19 (sklearn's
    LinearRegression restated as a least-squares solve with an optional trailing intercept column)."""
    X = np.hstack((C, np.ones((C.shape[0], 1)))) if fit_intercept else C
    coefs, *_ = np.linalg.lstsq(X, y, rcond=None)
    r = y - X.dot(coefs)
    n = mats[0].shape[0]
    off = [sp.csr_matrix(m - m.multiply(sp.eye(n))) for m in mats]
    m = len(mats)
    xtx = np.zeros((m, m))
    for i in range(m):
        for j in range(i, m):
            xtx[i, j] = xtx[j, i] = off[i].multiply(off[j]).sum()
    xty = np.array([r.dot(off[i].dot(r)) for i in range(m)])
    coef = np.linalg.inv(xtx).dot(xty)
    return np.append(coef, 1 - coef.sum()), coefs.tolist()
