"""Input generators (host): the vectorised IBD builder is bit-identical to the reference's LD/numerator on
the pedigrees frozen in the golden files; simulator and household matrix basic properties."""
import numpy as np
import scipy.sparse as sp

from scilmm_b200 import pedigree as P


def test_numerator_bit_exact_vs_reference(golden_small, golden_c1mini):
    for g in (golden_small, golden_c1mini):
        rel = g.csr("rel")
        A, T, D, F = P.numerator(rel)
        Ag, Tg = g.csr("ibd_full"), g.csr("ibd_Lfac")
        assert np.array_equal(A.indptr, Ag.indptr) and np.array_equal(A.indices, Ag.indices)
        assert np.array_equal(A.data, Ag.data)
        assert abs(T - Tg).max() == 0
        assert np.array_equal(D, g["ibd_D"])


def test_hand_checked_pedigree():
    # the reference's 10-individual fixture in topological order: known IBD entries
    par = {3: (0, 1), 4: (0, 1), 6: (2, 3), 7: (4, 5), 8: (6, 7), 9: (6, 7)}
    rows, cols = [], []
    for c, ps in par.items():
        for p in ps:
            rows.append(c)
            cols.append(p)
    rel = sp.csr_matrix((np.ones(len(rows), bool), (rows, cols)), shape=(10, 10))
    A, T, D, F = P.numerator(rel)
    A = A.toarray()
    assert A[3, 4] == 0.5 and A[6, 7] == 0.125 and A[8, 9] == 0.5625 and A[9, 9] == 1.0625
    assert np.allclose(D, [1, 1, 1, .5, .5, 1, .5, .5, .5, .5])


def test_simulator_and_household():
    ped = P.simulate_pedigree(3000, 0.003, seed=3)
    rel = ped["rel"]
    assert sp.triu(rel).nnz == 0 and rel.shape == (3000, 3000)
    assert np.diff(rel.indptr).max() <= 2
    A, T, D, F = P.numerator(rel)
    assert abs(A.nnz - 3000 ** 2 * 0.003) < 0.1 * 3000 ** 2 * 0.003
    assert abs(A - A.T).max() < 1e-15 and A.diagonal().min() >= 1.0
    H = P.household_matrix(ped["household"])
    assert abs(H - H.T).nnz == 0 and np.all(H.diagonal() == 1) and set(np.unique(H.data)) == {1.0}
    ped2 = P.simulate_pedigree(3000, 0.003, seed=3)
    assert (ped2["rel"] != rel).nnz == 0            # seeded
    keep, (Af, Hf) = P.drop_unrelated(A, H)
    assert Af.shape[0] == keep.sum() == Hf.shape[0]
    assert np.all(np.asarray(Af.sum(axis=1)).ravel() > 1)


def test_npz_csr_io_matches_reference_format(tmp_path, golden_small):
    """save_sparse_csr / load_sparse_csr (reference Matrices/SparseMatrixFunctions.py:5-13): same keys, exact round
    trip, and a file written the way the reference writes it loads identically."""
    import scilmm_b200
    A = golden_small.csr("A")
    scilmm_b200.save_sparse_csr(str(tmp_path / "IBD"), A)                 # np.savez appends .npz
    z = np.load(tmp_path / "IBD.npz")
    assert set(z.files) == {"data", "indices", "indptr", "shape"} and tuple(z["shape"]) == A.shape
    B = scilmm_b200.load_sparse_csr(str(tmp_path / "IBD.npz"))
    assert np.array_equal(B.indptr, A.indptr) and np.array_equal(B.indices, A.indices) and np.array_equal(B.data, A.data)
    np.savez(tmp_path / "ref.npz", data=A.data, indices=A.indices, indptr=A.indptr, shape=A.shape)   # reference :6-7
    Cm = scilmm_b200.load_sparse_csr(str(tmp_path / "ref.npz"))
    assert (Cm != A).nnz == 0
    E = scilmm_b200.pairwise_epistasis(A)
    assert np.array_equal(sp.csr_matrix(E).data, golden_small.csr("E").data)


def test_simulate_tree_is_stream_identical_to_the_reference(golden_small, golden_c1mini, golden_c1):
    """scilmm_b200.pedigree.simulate_tree draws from the legacy global numpy stream exactly like the reference's
    Simulation/Pedigree.py:93-124: with the seed the golden files were frozen with, the relationship matrix, and the
    phenotype of Simulation/Phenotype.py:24-35 drawn next, are the reference's bit for bit (config 1 included)."""
    for g, n_sim, sf in ((golden_small, 400, 0.01), (golden_c1mini, 2500, 0.004), (golden_c1, 10000, 0.001)):
        np.random.seed(g.seed)
        rel, sex, gen, hh = P.simulate_tree(n_sim, sf, 1.4, 0.8, return_households=True)
        rel = sp.csr_matrix(rel)
        rel.sort_indices()
        R = g.csr("rel")
        assert np.array_equal(rel.indptr, R.indptr) and np.array_equal(rel.indices, R.indices)
        assert sex.sum() > 0 and gen.max() >= 3 and hh.max() > 0 and np.all(hh[gen == 0] == -1)
        A, T, D, F = P.numerator(rel)
        np.random.seed(g.seed + 1)
        cov_raw = np.random.randn(n_sim, 2)
        y = P.quick_simulate_phenotype(T @ sp.diags(np.sqrt(D)), cov_raw, 0.4, np.arange(1, 3) * 0.1)
        assert np.array_equal(y[g["keep"]], g["y"])
