"""File formats and matrix helpers either side of the hot path, mirroring the reference's
scilmm/Matrices package name for name:

    save_sparse_csr / load_sparse_csr   Matrices/SparseMatrixFunctions.py:5-13  (.npz with data/indices/indptr/shape:
                                         the format IBDCompute.compute_ibd writes IBD.npz / L.npz / D.npz in, :82-84)
    pairwise_epistasis                  Matrices/Epistasis.py:1-2
    simple_numerator / create_numerator Matrices/Numerator.py:37-43 on the GPU builder (scilmm_b200.ibd) when a device is
                                         present - host numpy otherwise is NOT provided here: see scilmm_b200.pedigree
"""
import numpy as np
import scipy.sparse as sp


def save_sparse_csr(filename, array):
    """reference Matrices/SparseMatrixFunctions.py:5-7 (np.savez appends .npz when missing, as there)."""
    array = sp.csr_matrix(array)
    np.savez(filename, data=array.data, indices=array.indices, indptr=array.indptr, shape=array.shape)


def load_sparse_csr(filename):
    """reference Matrices/SparseMatrixFunctions.py:10-13."""
    loader = np.load(filename)
    return sp.csr_matrix((loader['data'], loader['indices'], loader['indptr']), shape=tuple(loader['shape']))


def pairwise_epistasis(ibd):
    """reference Matrices/Epistasis.py:1-2: the Hadamard square (same sparsity pattern as the IBD matrix)."""
    return ibd.multiply(ibd)
