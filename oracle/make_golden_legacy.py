"""TEST INFRASTRUCTURE ONLY - golden vectors for the reference's legacy entry points, produced by running the
UNMODIFIED reference modules scilmm/Estimation/LMM.py (LMM, :154-171) and scilmm/Estimation/HE.py (compute_HE,
:22-40) on the inputs of tests/golden/case_small.npz.

Run in the authoring container (needs /root/reference):   python oracle/make_golden_legacy.py
CHOLMOD is replaced by oracle.cpu_factor.DenseFactor (identity permutation), as in oracle/make_golden.py; the
probe stream is the legacy global numpy stream, seeded before the call.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.cpu_factor import DenseFactor  # noqa: E402
from oracle.refload import load_reference  # noqa: E402
from tests.util import load_golden  # noqa: E402


def main():
    load_reference()
    import scilmm.Estimation.LMM as RL
    import scilmm.Estimation.HE as RH
    g = load_golden("case_small")
    A, E, H = g.csr("A"), g.csr("E"), g.csr("H")
    cov_raw = g["cov"][:, :-1].copy()          # the golden covariates carry the intercept LAST; the legacy path adds it FIRST
    y = g["y"].copy()
    out = {}
    np.random.seed(21)
    res = RL.LMM(lambda V: DenseFactor(V), [A, E], cov_raw, y.copy(), with_intercept=True, reml=True, sim_num=20,
                 verbose=False)
    out["lmm_sig"] = res["covariance coefficients"]
    out["lmm_beta"] = res["covariates coefficients"]
    out["lmm_se"] = res["covariance std"]
    out["lmm_pvalues"] = res["covariates p-values"]
    for fit in (False, True):
        coef, cc = RH.compute_HE(y.copy(), cov_raw, [A, E, H], fit_intercept=fit)
        out["he_coef_%d" % fit] = coef
        out["he_covcoef_%d" % fit] = np.asarray(cc, dtype=np.float64)
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "case_small_legacy.npz")
    np.savez_compressed(path, **out)
    for k, v in out.items():
        print(k, v)


if __name__ == "__main__":
    main()
