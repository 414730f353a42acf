"""TEST INFRASTRUCTURE ONLY. Loads the *unmodified* reference from /root/reference (authoring container only).

Used by oracle/make_golden.py to freeze golden vectors and by tests that cross-check the oracle
restatement against the real reference when it is present.  Never imported by the product package.

Shims:
  1. stub `sksparse.cholmod` on sys.path (reference SparseCholesky.py:10),
  2. `np.float = float` (reference SparseCholesky.py:384),
  3. ragged np.array in Simulation/Pedigree.py:44,54 -> object arrays (monkeypatched `households`).
"""
import os
import sys

import numpy as np

REFERENCE_ROOT = os.environ.get("SCILMM_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "scilmm"))


def load_reference():
    """Return (SC, sim_pedigree, numerator, phenotype) reference modules."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    shim = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_shim")
    for p in (REFERENCE_ROOT, shim):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not hasattr(np, "float"):
        np.float = float
    import scilmm  # noqa: F401
    SC = sys.modules["scilmm.SparseCholesky"]
    import scilmm.Simulation.Pedigree as ped
    import scilmm.Matrices.Numerator as num
    import scilmm.Simulation.Phenotype as phe

    if not getattr(ped, "_oracle_patched", False):
        _orig_array = np.array

        class _NPProxy(object):
            """numpy proxy whose .array falls back to an object array for ragged input."""

            def __getattr__(self, name):
                return getattr(np, name)

            @staticmethod
            def array(obj, *a, **k):
                try:
                    return _orig_array(obj, *a, **k)
                except ValueError:
                    out = np.empty(len(obj), dtype=object)
                    for i, v in enumerate(obj):
                        out[i] = v
                    return out

        ped.np = _NPProxy()
        ped._oracle_patched = True
    return SC, ped, num, phe
