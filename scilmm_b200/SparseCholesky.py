"""Drop-in for the reference module scilmm/SparseCholesky.py on a B200.

Same names, argument meaning and return values as the reference:
    SparseCholesky (the `cholesky_func` functor, reference :16-26)  ->  Factor with __call__/logdet/L/P
    HE (:192-281), REML (:177-189), bolt_gradient_estimation (:77-117), estimate_var_comps (:120-144),
    compute_hess (:147-168), compute_varcomp_stderr (:171-174), matrices_weighted_sum (:55-59),
    estimate_fixed_effects (:29-34), negative_log_likelihood (:37-46), simulate_vector (:49-52),
    compute_gradients (:62-74), MINQUE (:284-347), run_estimates (:350-395).

All sparse / factorization arithmetic runs in libscilmm_b200.so on the GPU.  The symbolic analysis
(ordering, elimination tree, supernodes, schedules) is cached per sparsity pattern inside the functor,
which is the structural change against the reference (it re-analyses on every call, :22-26 from :92).
There is no CPU fallback: without the CUDA library these functions raise.
"""
import hashlib
import time

import numpy as np
import scipy.linalg as la
import scipy.optimize as optimize
import scipy.sparse as sparse

from . import engine as _eng
from . import sharding as _shard
from ._lib import NotPositiveDefiniteError, SlmmError  # noqa: F401  (re-exported)

LOG_2PI = np.log(2 * np.pi)


def _pattern_key(m):
    h = hashlib.blake2b(digest_size=16)
    h.update(np.ascontiguousarray(m.indptr).view(np.uint8))
    h.update(np.ascontiguousarray(m.indices).view(np.uint8))
    return (m.shape[0], int(m.nnz), h.hexdigest())


class B200Factor(object):
    """The Factor protocol of sksparse.cholmod (the four members the reference uses)."""

    def __init__(self, chol):
        self._chol = chol
        self._serial = chol.factorizations

    def _check(self):
        if self._chol.factorizations != self._serial:
            raise RuntimeError("this Factor was invalidated by a later factorization with the same pattern "
                               "(the device storage is reused across REML iterations)")

    # factor(b)  -- reference call sites :30,32,52,100,149,153
    def __call__(self, b):
        self._check()
        torch = _eng.require_cuda()
        if torch.is_tensor(b):
            x = b.to(device="cuda", dtype=torch.float64).contiguous().clone()
            return self._chol.solve_(x)
        b = np.asarray(b, dtype=np.float64)
        x = _eng.to_device(b, torch).contiguous()
        self._chol.solve_(x)
        return x.cpu().numpy()

    solve_A = __call__

    def logdet(self):
        self._check()
        return self._chol.logdet()

    def L(self):
        self._check()
        return self._chol.export_L()

    def P(self):
        return self._chol.perm().copy()

    # device-side extras used by the fused path
    def lmul(self, Z):
        self._check()
        return self._chol.lmul(Z)

    def stats(self):
        return self._chol.stats()


class SparseCholesky(object):
    """cholesky_func functor, reference scilmm/SparseCholesky.py:16-26.

    use_long / mode are accepted for signature compatibility (the engine is always supernodal LL' with
    32-bit row indices and 64-bit pointers).  ordering_method: 'nesdis' / 'metis' (nested dissection),
    'natural', or pass `perm` (perm[new] = old) to force a permutation (parity mode: L is unique given P).
    rng: 'numpy' draws probe vectors from the global numpy stream exactly like the reference (:50);
         'device' draws them on the GPU (distribution-equivalent, not stream-identical);
         'host_buffer' takes them from `probe_source(n, sim_num)` (a host array / pinned tensor).
    """

    def __init__(self, use_long=False, mode='supernodal', ordering_method='nesdis', perm=None, rng='numpy'):
        self._use_long = use_long
        self._mode = mode
        self._ordering_method = ordering_method
        self._perm = perm
        self.rng = rng
        self.probe_source = None      # callable(n, sim_num) -> host block, used when rng == 'host_buffer'
        self._engines = {}
        self._sessions = {}
        self.timings = {}

    def _engine_for(self, pattern):
        key = _pattern_key(pattern)
        eng = self._engines.get(key)
        if eng is None:
            t0 = time.time()
            eng = _eng.CholEngine(pattern, ordering=self._ordering_method, perm=self._perm)
            eng._self_map = eng.register_pattern(pattern)
            self.timings['analyze_s'] = time.time() - t0
            if len(self._engines) >= 4:
                self._engines.pop(next(iter(self._engines)))
            self._engines[key] = eng
        return eng

    def __call__(self, sparse_mat):
        torch = _eng.require_cuda()
        if not sparse.issparse(sparse_mat):
            raise TypeError("SparseCholesky expects a scipy.sparse matrix")
        m = sparse_mat
        if not sparse.isspmatrix_csc(m) and not sparse.isspmatrix_csr(m):
            m = m.tocsc()
        if m.shape[0] != m.shape[1]:
            raise ValueError("matrix must be square")
        if not m.has_sorted_indices:
            m = m.sorted_indices()
        pat = sparse.csr_matrix((m.data, m.indices, m.indptr), shape=m.shape)   # symmetric: CSC == CSR
        pat = _eng.canonical_csr(pat)
        eng = self._engine_for(pat)
        vals = _eng.to_device(pat.data, torch)
        eng.add_values(eng._self_map, vals.data_ptr(), 1.0, True)
        eng.factorize()
        return B200Factor(eng)

    # ---- fused session for the REML objective: matrices resident, patterns registered once
    def _session(self, mats, covariates, y):
        key = (tuple(id(m) for m in mats), id(covariates), id(y))
        ses = self._sessions.get(key)
        if ses is None:
            self._sessions.clear()
            ses = RemlSession(self, mats, covariates, y)
            self._sessions[key] = ses
        return ses


class RemlSession(object):
    """Device-resident state of one REML problem: K matrices, covariates, phenotype, one symbolic analysis."""

    def __init__(self, functor, mats, covariates, y):
        torch = _eng.require_cuda()
        self.torch = torch
        self.functor = functor
        self._refs = (list(mats), covariates, y)     # keeps the ids used as the cache key alive
        self.mats_host = [_eng.canonical_csr(m) for m in mats]
        self.K = len(mats)
        self.n = self.mats_host[0].shape[0]
        t0 = time.time()
        union = None
        for m in self.mats_host:        # pattern union (values are irrelevant; ones avoid cancellation)
            ones = sparse.csr_matrix((np.ones(m.nnz), m.indices, m.indptr), shape=m.shape)
            union = ones if union is None else union + ones
        union = _eng.canonical_csr(union)
        self.union = union
        self.eng = functor._engine_for(union)
        self.matset = _eng.MatSet(self.mats_host)
        self.map_ids = []
        cache = {}
        for m in self.mats_host:
            k = _pattern_key(m)
            if k not in cache:
                cache[k] = self.eng.register_pattern(m)
            self.map_ids.append(cache[k])
        self.C = _eng.to_device(np.asarray(covariates, dtype=np.float64), torch)
        self.y = _eng.to_device(np.asarray(y, dtype=np.float64), torch)
        self.C_host = np.asarray(covariates, dtype=np.float64)
        self.y_host = np.asarray(y, dtype=np.float64)
        self.setup_s = time.time() - t0
        self._groups = None
        self.last = {}

    # K1 + K2
    def factor_at(self, sigmas):
        k = 0
        while k < self.K:      # consecutive matrices with one pattern (A, A o A) are scattered in one pass
            if k + 1 < self.K and self.map_ids[k + 1] == self.map_ids[k]:
                self.eng.add_values2(self.map_ids[k], self.matset.values_ptr(k), float(sigmas[k]),
                                     self.matset.values_ptr(k + 1), float(sigmas[k + 1]), k == 0)
                k += 2
            else:
                self.eng.add_values(self.map_ids[k], self.matset.values_ptr(k), float(sigmas[k]), k == 0)
                k += 1
        self.eng.factorize()
        return B200Factor(self.eng)

    def fixed_effects(self):
        """V^-1 C, chol(C'V^-1C), mu, beta, V^-1 y   (reference :29-34, one multi-RHS solve)."""
        return self._fixed_effects_finish(self._fixed_effects_start(overlap=False))

    def _fixed_effects_start(self, overlap=True):
        """Queues the (c+1)-column solve.  With overlap=True it goes to the engine's auxiliary stream: the narrow
        solve is a launch-latency-bound chain that leaves the SMs idle, so it runs beside the probe pipeline the
        caller issues on stream 0 next; _fixed_effects_finish joins the two."""
        torch = self.torch
        B = torch.cat([self.C, self.y.unsqueeze(1)], dim=1).contiguous()
        if overlap:
            self.eng.aux_begin()
            try:
                self.eng.solve_(B)
            finally:
                self.eng.aux_end()
        else:
            self.eng.solve_(B)
        return B, overlap

    def _fixed_effects_finish(self, pending):
        B, overlap = pending
        if overlap:
            self.eng.aux_join()
        c = self.C.shape[1]
        self._ViCy = B                                  # [V^-1 C | V^-1 y], kept for the Gram pass
        ViC = B[:, :c].contiguous()
        Viy = B[:, c].contiguous()
        CtViC = (self.C.t() @ ViC).cpu().numpy()
        CtViy = (self.C.t() @ Viy).cpu().numpy()
        chol = la.cho_factor(CtViC)
        beta = la.cho_solve(chol, CtViy)
        return ViC, chol, beta, Viy

    def probes(self, sim_num, Z=None, col_begin=0, col_end=None):
        """W = V^-1 (L Z)[argsort P]   (reference :49-52)."""
        torch = self.torch
        if Z is None:
            if self.functor.rng == 'numpy':
                Z = _eng.to_device(np.random.randn(self.n, sim_num), torch)
            elif self.functor.rng == 'host_buffer':       # caller-supplied host block (pinned tensor or ndarray)
                Z = self.functor.probe_source(self.n, sim_num)
            else:
                Z = torch.randn(self.n, sim_num, dtype=torch.float64, device="cuda")
        if not torch.is_tensor(Z):
            Z = _eng.to_device(np.asarray(Z, dtype=np.float64), torch)
        elif not Z.is_cuda:
            Z = Z.to("cuda", non_blocking=True)
        if col_end is not None or col_begin:
            Z = Z[:, col_begin:col_end].contiguous()
        U = self.eng.lmul(Z)
        return self.eng.solve_(U)

    def evaluate(self, sigmas, reml, sim_num, Z=None):
        """One REML evaluation: nll and d nll / d sigma  (reference :77-117 without the exp chain rule)."""
        torch = self.torch
        self.factor_at(sigmas)
        logdet = self.eng.logdet()
        n = self.n
        pending = self._fixed_effects_start()          # narrow solve on the auxiliary stream ...
        # ... beside the probe pipeline; probe columns are sharded across ranks when torch.distributed is initialised
        rank, world = _shard.rank_world()
        if Z is None and self.functor.rng == 'numpy' and world > 1:
            Z = np.random.randn(n, sim_num)            # every rank draws the same stream, keeps its slice
        lo, hi = _shard.column_block(sim_num, rank, world)
        W = self.probes(sim_num, Z, lo, hi) if world > 1 else self.probes(sim_num, Z)
        ViC, chol, beta, Viy = self._fixed_effects_finish(pending)
        beta_t = torch.from_numpy(beta).to("cuda")
        Vir = Viy - ViC @ beta_t                       # V^-1 (y - C beta) by linearity
        r = self.y - self.C @ beta_t
        nll = 0.5 * (float(r @ Vir) + n * LOG_2PI + logdet)
        if reml:
            nll += 0.5 * 2 * np.sum(np.log(np.diag(chol[0])))
        c = self.C.shape[1]
        K = self.K
        comp1 = torch.zeros(K, dtype=torch.float64, device="cuda")
        comp2 = torch.zeros(K, dtype=torch.float64, device="cuda")
        gram = torch.zeros(K, c, c, dtype=torch.float64, device="cuda")
        # Per pattern group ONE pass over the matrices: column quadratic forms of the probes W (symmetric matrices
        # are traversed on/below the diagonal only) and, riding along, the Gram matrix of the narrow block
        # B = [V^-1 C | V^-1 r]: its last diagonal entry is r'V^-1 A_k V^-1 r (:66), its leading c x c block feeds
        # the REML trace term (:70).
        s_loc = W.shape[1]
        if self._groups is None:
            self._groups = self.matset.pattern_groups(2)
            self._sym = [self.matset.is_symmetric(k) for k in range(K)]
        Bq = self._ViCy
        Bq[:, c] = Vir                                 # Viy was copied out above; the block becomes [V^-1 C | V^-1 r]
        for ks in self._groups:
            if all(self._sym[k] for k in ks) and Bq.shape[1] <= 16 and s_loc <= 160:
                dots, G = self.matset.quadform_gram_multi(ks, W, Bq)
                for g, k in enumerate(ks):
                    comp1[k] = dots[g].sum()
                    comp2[k] = G[g][c, c]
                    gram[k] = G[g][:c, :c]
            else:                                      # general matrices / very wide blocks: separate passes
                X = torch.cat([W, Vir.unsqueeze(1)], dim=1).contiguous()
                dots = self.matset.quadform_multi(ks, X)
                if reml:
                    _, stored = self.matset.coldot_multi(ks, ViC, 0)
                for g, k in enumerate(ks):
                    comp1[k] = dots[g, :s_loc].sum()
                    comp2[k] = dots[g, s_loc]
                    if reml:
                        gram[k] = ViC.t() @ stored[g]
        _shard.allreduce_sum_(comp1)                   # the only collective of an evaluation: K doubles
        comp1 = (comp1 / sim_num).cpu().numpy()
        comp2 = comp2.cpu().numpy()
        grad = 0.5 * (comp1 - comp2)
        if reml:
            gram_h = gram.cpu().numpy()
            for k in range(K):
                grad[k] -= 0.5 * np.trace(la.cho_solve(chol, gram_h[k]))
        self.last = dict(logdet=logdet, beta=beta, ViC=ViC, Vir=Vir, Viy=Viy, W=W, chol=chol)
        return nll, grad


def _need_functor(cholesky_func):
    if not isinstance(cholesky_func, SparseCholesky):
        raise TypeError("scilmm_b200 needs its own SparseCholesky() functor as cholesky_func "
                        "(the engine has no CPU / CHOLMOD path)")
    return cholesky_func


# ------------------------------------------------------------------------------------------- reference API
def matrices_weighted_sum(mats, sig2g_array):
    """reference :55-59.  Host scipy assembly for callers that want V itself; the fused path never forms it."""
    V = sig2g_array[0] * mats[0]
    for i in range(1, len(sig2g_array)):
        V = V + sig2g_array[i] * mats[i]
    return V.tocsc()


def estimate_fixed_effects(factor, y, covariates):
    """reference :29-34."""
    invV_C = factor(covariates)
    L_CT_invV_C = la.cho_factor(covariates.T.dot(invV_C))
    fixed_effects = la.cho_solve(L_CT_invV_C, covariates.T.dot(factor(y)))
    mu = covariates.dot(fixed_effects)
    return invV_C, L_CT_invV_C, mu, fixed_effects


def negative_log_likelihood(factor, y, invV_y, mu, L_CT_invV_C, reml):
    """reference :37-46."""
    n = y.size
    nll = 0.5 * ((y - mu).dot(invV_y) + n * LOG_2PI + factor.logdet())
    if reml:
        nll += 0.5 * 2 * np.sum(np.log(np.diag(L_CT_invV_C[0])))
    return nll


def simulate_vector(factor, n, sim_num, p_inv=None):
    """reference :49-52: V^-1 (L Z)[argsort P] with Z from the global numpy stream; stays on the device
    until the final copy.  p_inv is accepted for signature compatibility (the engine applies P itself)."""
    torch = _eng.require_cuda()
    Z = _eng.to_device(np.random.randn(n, sim_num), torch)
    U = factor.lmul(Z)
    return factor._chol.solve_(U).cpu().numpy()


def compute_gradients(sig2g_array, mats, sim_vec, invV_y, reml, invV_C, L_CT_invV_C):
    """reference :62-74 on the GPU kernels (fused SpMM + column reductions)."""
    torch = _eng.require_cuda()
    ms = _eng.MatSet(mats)
    W = _eng.to_device(np.asarray(sim_vec, dtype=np.float64), torch)
    r = _eng.to_device(np.asarray(invV_y, dtype=np.float64), torch)
    X = torch.cat([W, r.unsqueeze(1)], dim=1).contiguous()
    ViC = _eng.to_device(np.asarray(invV_C, dtype=np.float64), torch)
    grad = np.zeros(len(sig2g_array))
    for k in range(len(sig2g_array)):
        d = ms.quadform_multi([k], X)[0].cpu().numpy()
        grad[k] = 0.5 * (np.mean(d[:-1]) - d[-1])
        if reml:
            vec = (ViC.t() @ ms.spmm(k, ViC)).cpu().numpy()
            grad[k] -= 0.5 * np.trace(la.cho_solve(L_CT_invV_C, vec))
    return grad


def bolt_gradient_estimation(log_sig2g_array, cholesky_func, mats, covariates, y, reml, sim_num, verbose,
                             take_exp=True):
    """One REML evaluation (nll, gradient), reference :77-117."""
    functor = _need_functor(cholesky_func)
    sig2g_array = np.exp(log_sig2g_array) if take_exp else np.asarray(log_sig2g_array, dtype=float)
    if verbose:
        t0 = time.time()
        print('estimating nll and its gradient at:', sig2g_array)
    ses = functor._session(mats, covariates, y)
    nll, grad = ses.evaluate(sig2g_array, reml, sim_num)
    if take_exp:
        grad = grad * sig2g_array
    if verbose:
        print("grad : ", grad)
        print('nll: %0.8e   computation time: %0.2f seconds' % (nll, time.time() - t0))
    return nll, grad


def estimate_var_comps(cholesky_func, mats, covariates, y, reml=True, sim_num=100, verbose=True, aireml=False):
    """reference :120-144."""
    he_est = HE(mats[:-1], covariates, y, compute_stderr=False)
    he_est = np.concatenate((he_est, [1 - he_est.sum()]))
    x0 = he_est
    if np.any(x0 < 0):
        x0 = np.ones((len(mats)))
    x0 = x0 / x0.sum()
    if aireml:
        raise NotImplementedError('AI-REML is broken')
    optObj = optimize.minimize(bolt_gradient_estimation, np.log(x0),
                               args=(cholesky_func, mats, covariates, y, reml, sim_num, verbose, True),
                               jac=True, method='L-BFGS-B', options={'eps': 1e-5, 'ftol': 1e-7})
    if not optObj.success:
        print('optimization failed with message: %s' % optObj.message)
    return np.exp(optObj.x)


def _hess_device(ses, factor_eng):
    """-0.5 y' P A_i P A_j P y on the device (reference :147-168)."""
    torch = ses.torch
    K = ses.K
    C, y = ses.C, ses.y
    ViC = factor_eng.solve_(C.clone().contiguous())
    chol = la.cho_factor((C.t() @ ViC).cpu().numpy())
    Minv = torch.from_numpy(la.cho_solve(chol, np.eye(C.shape[1]))).to("cuda")

    def project(Zb):          # Zb: n x k block
        Viz = factor_eng.solve_(Zb.clone().contiguous())
        return Viz - ViC @ (Minv @ (C.t() @ Viz))

    Py = project(y.unsqueeze(1))
    AjPy = torch.cat([ses.matset.spmm(j, Py) for j in range(K)], dim=1)       # n x K
    PAjPy = project(AjPy)
    hess = np.empty((K, K))
    for j in range(K):
        cols = torch.cat([ses.matset.spmm(i, PAjPy[:, j:j + 1].contiguous()) for i in range(j + 1)], dim=1)
        Pcols = project(cols)
        vals = (-0.5 * (y @ Pcols)).cpu().numpy()
        for i in range(j + 1):
            hess[i, j] = vals[i]
            hess[j, i] = vals[i]
    return hess


def compute_hess(mats, covariates, factor, y):
    """reference :147-168.  `factor` must be a B200Factor; solves are batched per stage."""
    if not isinstance(factor, B200Factor):
        raise TypeError("compute_hess needs a B200Factor")
    factor._check()
    ses = _AdHocSession(mats, covariates, y)
    return _hess_device(ses, factor._chol)


class _AdHocSession(object):
    def __init__(self, mats, covariates, y):
        torch = _eng.require_cuda()
        self.torch = torch
        self.K = len(mats)
        self.matset = _eng.MatSet(mats)
        self.C = _eng.to_device(np.asarray(covariates, dtype=np.float64), torch)
        self.y = _eng.to_device(np.asarray(y, dtype=np.float64), torch)


def compute_varcomp_stderr(mats, covariates, factor, y, sim_num):
    """reference :171-174."""
    hess = compute_hess(mats, covariates, factor, y)
    inv_neg_hess = la.inv(-hess)
    return np.sqrt(np.diag(inv_neg_hess) * (1 + 1.0 / sim_num))


def REML(cholesky_func, mats, covariates, y, reml=True, sim_num=100, verbose=False):
    """reference :177-189.  Returns the reference's three keys plus 'nll' (the log-likelihood the reference
    computes at :103 but drops) and 'engine' statistics."""
    functor = _need_functor(cholesky_func)
    y = y / y.std()
    mats = mats + [sparse.eye(y.shape[0]).tocsr()]
    varcomp_estimates = estimate_var_comps(functor, mats, covariates, y, reml, sim_num, verbose)
    ses = functor._session(mats, covariates, y)
    factor = ses.factor_at(varcomp_estimates)
    ViC, chol, fixed_effects, Viy = ses.fixed_effects()
    torch = ses.torch
    beta_t = torch.from_numpy(fixed_effects).to("cuda")
    Vir = Viy - ViC @ beta_t
    r = ses.y - ses.C @ beta_t
    nll = 0.5 * (float(r @ Vir) + y.size * LOG_2PI + factor.logdet())
    if reml:
        nll += 0.5 * 2 * np.sum(np.log(np.diag(chol[0])))
    hess = _hess_device(ses, ses.eng)
    sigmas_sigmas = np.sqrt(np.diag(la.inv(-hess)) * (1 + 1.0 / sim_num))
    return {"covariance coefficients": varcomp_estimates,
            "covariates coefficients": fixed_effects,
            "covariance std": sigmas_sigmas,
            "nll": nll,
            "engine": ses.eng.stats()}


# ------------------------------------------------------------------------------------------- Haseman-Elston
def he_moments(mat_list, y, MQS=False, matset=None):
    """q and S of the HE normal equations on the GPU (reference :213-243)."""
    ms = matset if matset is not None else _eng.MatSet(mat_list)
    torch = _eng.require_cuda()
    K, n = ms.K, ms.n
    y_dev = _eng.to_device(np.asarray(y, dtype=np.float64), torch)
    rank, world = _shard.rank_world()
    if world == 1:
        out = ms.he_moments_device(y_dev).cpu().numpy()
    else:       # row blocks balanced by nonzeros + one all-reduce of 2K + 2K^2 doubles
        big = max(range(K), key=lambda k: ms.nnz[k])
        bounds = _shard.row_blocks_by_nnz(_eng._csr_arrays(mat_list[big])[0], world)
        part = ms.he_moments_device(y_dev, int(bounds[rank]), int(bounds[rank + 1])).clone()
        out = _shard.allreduce_sum_(part).cpu().numpy()
    q_off, q_diag, S_off, S_diag = _eng.MatSet.split_moments(out, K)
    if MQS:
        yy = float(np.dot(y, y))
        q = (q_off + q_diag) - yy
        S = (S_off + S_diag) - (n - 1)
    else:
        q, S = q_off, S_off
    return q, S, ms


def HE(mat_list, cov, y, MQS=False, verbose=False, sim_num=100, compute_stderr=False, y2=None):
    """Haseman-Elston regression, reference :192-281 (same outputs, including the bivariate mode's in-place
    mutation of mat_list / y2 and the reference's K>1 indexing quirks in the sampling-variance branch)."""
    if any(not sparse.issparse(m) for m in mat_list):
        raise TypeError("scilmm_b200.HE takes scipy.sparse matrices (the dense branch of the reference is CPU-only)")
    CTC = cov.T.dot(cov)
    y = y - cov.dot(np.linalg.solve(CTC, cov.T.dot(y)))
    y /= y.std()
    if y2 is not None:
        y2 -= cov.dot(np.linalg.solve(CTC, cov.T.dot(y2)))
        y2 /= y2.std()
        y = np.concatenate((y, y2))
        for m_i, m in enumerate(mat_list):
            z = sparse.csr_matrix((m.shape[0], m.shape[0]))
            mat_list[m_i] = sparse.vstack([sparse.hstack([z, m]), sparse.hstack([m, z])]).tocsr()
    K = len(mat_list)
    n = y.shape[0]
    q, S, ms = he_moments(mat_list, y, MQS)
    he_est = np.linalg.solve(S, q)
    if not compute_stderr:
        return he_est

    torch = _eng.require_cuda()
    stale_i = K - 1            # reference :254 reads the loop variable left over from :216
    stale_j = K - 1            # reference :272 reads mat_j left over from :224
    w = np.zeros(K)
    w[0] += he_est[0]
    for k in range(1, K):
        w[k] += he_est[stale_i]
    w_eye = 1.0 - he_est.sum()

    def Hdot(X):
        out = w_eye * X
        for k in range(K):
            if w[k] != 0.0:
                out = out + w[k] * ms.spmm(k, X)
        return out

    V_q = np.empty((K, K))
    if sim_num is None:
        raise NotImplementedError("the exact (sim_num=None) branch forms sparse x sparse products (reference "
                                  ":261,267) and has no GPU kernel; use sim_num")
    for i in range(K):
        for j in range(i + 1):
            Zs = _eng.to_device(np.random.randn(n, sim_num), torch)
            t1 = ms.spmm(stale_j, Zs) - Zs
            t2 = Hdot(t1)
            t3 = ms.spmm(j, t2) - t2          # the inner loop rebinds mat_i to mat_list[j] (:262)
            t4 = Hdot(t3)
            V_q[i, j] = 2 * float((Zs * t4).sum(dim=0).mean())
            V_q[j, i] = V_q[i, j]
    var_he_est = np.linalg.solve(S, np.linalg.solve(S, V_q).T).T
    return he_est, np.sqrt(np.diag(var_he_est))


def MINQUE(cholesky_func, mat_list, cov, y, compute_stderr=False, verbose=False, num_iter=100, sim_num=100):
    """Iterated MINQUE, reference :284-347 (same update rule, including the stale index at :335)."""
    functor = _need_functor(cholesky_func)
    torch = _eng.require_cuda()
    CTC = cov.T.dot(cov)
    y = y - cov.dot(np.linalg.solve(CTC, cov.T.dot(y)))
    y /= y.std()
    K = len(mat_list)
    n = y.shape[0]
    ms = _eng.MatSet(mat_list)
    y_dev = _eng.to_device(y, torch)
    yy = float(y.dot(y))
    H = None
    q = np.zeros(K)
    S = np.zeros((K, K))
    for iter_num in range(num_iter):
        q = np.zeros(K)
        S = np.zeros((K, K))
        if H is None:
            q0, S0, _ = he_moments(mat_list, y, MQS=True, matset=ms)
            q, S = q0, S0
        else:
            factor = functor(H)
            Z = _eng.to_device(np.random.randn(n, sim_num), torch)
            invH_simy = factor._chol.solve_(factor.lmul(Z))
            invH_y = factor._chol.solve_(y_dev.clone())
            for i in range(K):
                q[i] = float(invH_y @ ms.spmm(i, invH_y)) - yy
                t = factor._chol.solve_(ms.spmm(i, invH_simy).contiguous())
                for j in range(i + 1):
                    S[i, j] = float((invH_simy * ms.spmm(j, t)).sum(dim=0).mean()) - (n - 1)
                    S[j, i] = S[i, j]
        minque_est = np.linalg.solve(S, q)
        if verbose:
            print(iter_num + 1, minque_est)
        stale_i = K - 1
        H = mat_list[0] * minque_est[0]
        for m in mat_list[1:]:
            H = H + m * minque_est[stale_i]
        H = H + sparse.eye(n, format='csr') * (1.0 - minque_est.sum())
    minque_est = np.linalg.solve(S, q)
    if not compute_stderr:
        return minque_est
    return minque_est, np.sqrt(0)


def run_estimates(A, df_phe, df_cov, reml=False, ignore_indices=False, df_phe2=None):
    """reference :350-395: align by IID, drop individuals without relatives, standardise covariates, fit.
    Differences: np.float -> float (:384) and the REML branch returns the dict instead of raising at :390."""
    import pandas as pd
    if not ignore_indices:
        indices = list(set(df_cov.index) & set(df_phe.index) & set(A.indices))
        df_cov = df_cov.loc[indices]
        df_phe = df_phe.loc[indices]
        if df_phe2 is not None:
            df_phe2 = df_phe2.loc[indices]
        A = A[indices][:, indices]
    has_relatives = np.asarray(A.sum(axis=1))[:, 0] > 1
    if any(~has_relatives):
        A = A[has_relatives][:, has_relatives]
        df_cov = df_cov.loc[has_relatives]
        df_phe = df_phe.loc[has_relatives]
        if df_phe2 is not None:
            df_phe2 = df_phe2.loc[has_relatives]
    A.eliminate_zeros()
    y = np.asarray(df_phe.values, dtype=float).reshape(-1)
    if df_phe2 is not None:
        y2 = np.asarray(df_phe2.values, dtype=float).reshape(-1)
        assert not reml
    else:
        y2 = None
    if isinstance(df_cov, pd.Series):
        df_cov = df_cov.to_frame()
    df_cov = df_cov.copy()
    df_cov['intercept'] = 1
    cov = df_cov.values.copy().astype(float)
    cov[:, :-1] -= cov[:, :-1].mean(axis=0)
    cov[:, :-1] /= cov[:, :-1].std(axis=0)
    if reml:
        reml_d = REML(SparseCholesky(), [A], cov, y, verbose=True)
        print("reml d are %s and %s" % (reml_d["covariance coefficients"], reml_d["covariance std"]))
        return reml_d
    he_est = HE([A], cov, y, compute_stderr=True, y2=y2)
    print("HE estimates are %s and %s" % (he_est[0], he_est[1]))
    return he_est


def run_estimates_from_paths(A, phe, cov, reml=False, ignore_indices=False):
    """reference :398-403."""
    import pandas as pd
    from scipy.io import mmread
    A = mmread(A).tocsr()
    index_col = None if ignore_indices else 0
    df_cov = pd.read_csv(cov, index_col=index_col)
    df_phe = pd.read_csv(phe, header=None, index_col=index_col)
    return run_estimates(A, df_phe, df_cov, reml=reml, ignore_indices=ignore_indices)
