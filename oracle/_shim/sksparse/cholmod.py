"""Stub of sksparse.cholmod (scikit-sparse is not installed in this image).

The reference only needs the *name* `cholesky` at import time (scilmm/SparseCholesky.py:10);
every oracle run passes its own `cholesky_func`, so this is never called.
"""


class CholmodError(Exception):
    pass


class CholmodNotPositiveDefiniteError(CholmodError):
    pass


def cholesky(*args, **kwargs):
    raise RuntimeError("CHOLMOD is not available in this image; pass an oracle cholesky_func instead")
