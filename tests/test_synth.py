"""Synthetic matrices: closed-form row offsets of the C library vs the numpy twins, structural
properties (sorted, duplicate free, symmetric Laplacians) and torch/numpy R-MAT agreement."""
import numpy as np
import pytest

from sparsematrixvectormultiplication_b200 import synth


@pytest.mark.parametrize("n", [1, 2, 3, 7, 16])
def test_lap2d_offsets_and_structure(n, port):
    rp, ci, va = synth.lap2d_csr(n)
    offs = np.array([synth.row_offset(synth.SYNTH_LAP2D, n, row=r) for r in range(n * n + 1)])
    assert np.array_equal(offs, rp)
    y = port.spmv_csr_serial(rp, ci, va, np.ones(n * n))
    assert np.array_equal(y, 4.0 - np.diff(rp) + 1)  # 4 - (#neighbours): exact small integers
    for r in range(n * n):
        row = ci[rp[r]:rp[r + 1]]
        assert (np.diff(row) > 0).all()


@pytest.mark.parametrize("n", [1, 2, 3, 6])
def test_lap3d_offsets_and_structure(n, port):
    rp, ci, va = synth.lap3d_csr(n)
    offs = np.array([synth.row_offset(synth.SYNTH_LAP3D, n, row=r) for r in range(n ** 3 + 1)])
    assert np.array_equal(offs, rp)
    y = port.spmv_csr_serial(rp, ci, va, np.ones(n ** 3))
    assert np.array_equal(y, 6.0 - np.diff(rp) + 1)


def test_row_slices_are_consistent():
    n = 9
    rp, ci, va = synth.lap3d_csr(n)
    lo, hi = 100, 555
    rp2, ci2, va2 = synth.lap3d_csr(n, lo, hi)
    assert np.array_equal(rp2, rp[lo:hi + 1] - rp[lo])
    assert np.array_equal(ci2, ci[rp[lo]:rp[hi]]) and np.array_equal(va2, va[rp[lo]:rp[hi]])
    rp, ci, va = synth.uniform_csr(300, 640, 32)
    rp2, ci2, va2 = synth.uniform_csr(300, 640, 32, row_begin=17, row_end=200)
    assert np.array_equal(ci2, ci[rp[17]:rp[200]]) and np.array_equal(va2, va[rp[17]:rp[200]])


def test_uniform_is_sorted_stratified_and_in_unit_interval():
    M, N, k = 500, 4096, 32
    rp, ci, va = synth.uniform_csr(M, N, k)
    cols = ci.reshape(M, k)
    strata = cols // (N // k)
    assert (strata == np.arange(k)).all() and (np.diff(cols, axis=1) > 0).all()
    assert va.min() > 0.0 and va.max() <= 1.0
    assert synth.row_offset(synth.SYNTH_UNIFORM, M, N, k, 77) == 77 * k
    x = synth.hash_vector(1000)
    assert x.min() > 0.0 and x.max() <= 1.0 and len(np.unique(x)) == 1000


def test_rmat_torch_twin_matches_numpy_on_cpu():
    a = synth.rmat_csr(9, 8, seed=7, plant_dense_row=100)
    b = synth.rmat_csr_device(9, 8, seed=7, plant_dense_row=100, device="cpu")
    assert all(np.array_equal(u, v.numpy()) for u, v in zip(a, b))
    rp, ci, va = a
    assert np.diff(rp).max() >= 100
    for r in range(0, 512, 37):
        assert (np.diff(ci[rp[r]:rp[r + 1]]) > 0).all()  # sorted, duplicate free
