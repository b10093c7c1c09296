"""Drop-in host API (libspmv_b200.so) vs the reference: golden vectors + the oracle on random COO.
Bit-exact for every index and value array."""
import json

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from conftest import GOLDEN, golden_names, load_golden, random_coo
from sparsematrixvectormultiplication_b200 import host

THREADS = (1, 2, 3, 4, 8, 40)


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


@pytest.mark.parametrize("name", golden_names())
def test_parser_and_builders_match_golden(name):
    g = load_golden(name)
    pre = host.read_matrix_market(GOLDEN / "mtx" / f"{name}.mtx")
    assert (pre.M, pre.N, pre.nz) == (int(g["M"]), int(g["N"]), int(g["nz"]))
    assert pre.type == bytes(g["type"]).decode()
    assert np.array_equal(pre.I, g["I"]) and np.array_equal(pre.J, g["J"])
    assert np.array_equal(bits(pre.val), bits(g["val"]))
    csr = host.convert_in_csr(pre, name)
    assert (csr.M, csr.N, csr.nz) == (pre.M, pre.N, pre.nz)
    assert np.array_equal(csr.row_ptr, g["row_ptr"]) and np.array_equal(csr.col_idx, g["col_idx"])
    assert np.array_equal(bits(csr.values), bits(g["values"]))
    hll = host.convert_to_hll(pre)
    rows, maxnz, offset, JA, AS = hll.flat()
    assert np.array_equal(rows, g["hll_rows"]) and np.array_equal(maxnz, g["hll_maxnz"])
    assert np.array_equal(JA, g["hll_JA"]) and np.array_equal(bits(AS), bits(g["hll_AS"]))
    for b in range(hll.num_blocks):  # empty block: NULL arrays like the reference
        blk = hll.c.blocks[b]
        assert blk.N == pre.N
        assert (blk.MAXNZ == 0) == (not blk.JA) == (not blk.AS)
    for T in THREADS:
        s, e = host.prepare_thread_distribution(csr.M, csr.row_ptr, T, csr.nz)
        assert np.array_equal(np.stack([s, e]).reshape(2, -1), g[f"part_rows_{T}"]), T
        s, e = host.prepare_thread_distribution_hll(hll, T)
        assert np.array_equal(np.stack([s, e]).reshape(2, -1), g[f"part_hll_{T}"]), T


def test_parser_rejects_what_the_reference_rejects(capfd):
    for name in json.loads((GOLDEN / "errors.json").read_text()):
        with pytest.raises(host.HostApiError):
            host.read_matrix_market(GOLDEN / "mtx" / name)
    with pytest.raises(host.HostApiError):
        host.read_matrix_market(GOLDEN / "mtx" / "does_not_exist.mtx")


@settings(max_examples=120, deadline=None)
@given(st.integers(1, 140), st.integers(1, 70), st.integers(0, 700), st.booleans(), st.integers(0, 2**31))
def test_builders_match_oracle_on_random_coo(M, N, nz, dup, seed):
    from oracle import oracle as O
    chk = O.best_available()
    rng = np.random.default_rng(seed)
    coo = random_coo(rng, M, N, nz, dup=dup)
    pre = host.PreMatrix(coo.M, coo.N, coo.I, coo.J, coo.val)
    csr = host.convert_in_csr(pre)
    rp, ci, va = chk.coo_to_csr(coo)
    assert np.array_equal(csr.row_ptr, rp) and np.array_equal(csr.col_idx, ci) and np.array_equal(bits(csr.values), bits(va))
    hll = host.convert_to_hll(pre)
    h = chk.coo_to_hll(coo)
    rows, maxnz, offset, JA, AS = hll.flat()
    assert np.array_equal(rows, h.rows) and np.array_equal(maxnz, h.maxnz)
    assert np.array_equal(JA, h.JA) and np.array_equal(bits(AS), bits(h.AS))
    for T in (1, 2, 7, 33):
        s, e = host.prepare_thread_distribution(csr.M, csr.row_ptr, T, csr.nz)
        so, eo = chk.partition_rows(rp, T)
        assert np.array_equal(s, so) and np.array_equal(e, eo)
        s, e = host.prepare_thread_distribution_hll(hll, T)
        so, eo = chk.partition_hll(h, N, T)
        assert np.array_equal(s, so) and np.array_equal(e, eo)


def test_long_rows_with_and_without_duplicates(checker):
    """Rows above the builder's short-row threshold: merge-sort fast path (unique) and the
    reference-equivalent quicksort (duplicates), incl. sorted / reversed input."""
    rng = np.random.default_rng(11)
    for case in range(12):
        M, N, nz = 5, 4000, 3000
        coo = random_coo(rng, M, N, nz, dup=bool(case % 2))
        if case % 3 == 0:    # already sorted by (row, col)
            o = np.lexsort((coo.J, coo.I))
        elif case % 3 == 1:  # reverse sorted
            o = np.lexsort((coo.J, coo.I))[::-1]
        else:
            o = np.arange(coo.nz)
        coo.I, coo.J, coo.val = coo.I[o].copy(), coo.J[o].copy(), coo.val[o].copy()
        pre = host.PreMatrix(M, N, coo.I, coo.J, coo.val)
        csr = host.convert_in_csr(pre)
        rp, ci, va = checker.coo_to_csr(coo)
        assert np.array_equal(csr.col_idx, ci) and np.array_equal(bits(csr.values), bits(va)), case
        h = checker.coo_to_hll(coo)
        rows, maxnz, offset, JA, AS = host.convert_to_hll(pre).flat()
        assert np.array_equal(JA, h.JA) and np.array_equal(bits(AS), bits(h.AS)), case


def test_sort_row_is_the_reference_quicksort(reference):
    rng = np.random.default_rng(5)
    for n in (2, 3, 17, 200):
        c = rng.integers(0, 6, n).astype(np.int32)  # many duplicates
        v = rng.standard_normal(n)
        c1, v1 = c.copy(), v.copy()
        host.sort_row(c1, v1, 0, n - 1)
        c2, v2 = c.copy(), v.copy()
        reference.lib.sort_row.argtypes = [host.N.c_int_p, host.N.c_dbl_p, host.C.c_size_t, host.C.c_size_t]
        reference.lib.sort_row.restype = None
        reference.lib.sort_row(c2.ctypes.data_as(host.N.c_int_p), v2.ctypes.data_as(host.N.c_dbl_p), 0, n - 1)
        assert np.array_equal(c1, c2) and np.array_equal(bits(v1), bits(v2))


def test_partitioner_edge_cases(checker):
    rp = np.zeros(6, np.int32)  # five empty rows
    s, e = host.prepare_thread_distribution(5, rp, 3, 0)
    so, eo = checker.partition_rows(rp, 3)
    assert len(s) == len(so) == 0
    assert host.prepare_thread_distribution(0, np.zeros(1, np.int32), 4, 0)[0].size == 0
    rp = np.array([0, 10, 10, 10, 11], np.int32)  # trailing light rows
    for T in (1, 2, 3, 4, 9):
        s, e = host.prepare_thread_distribution(4, rp, T, 11)
        so, eo = checker.partition_rows(rp, T)
        assert np.array_equal(s, so) and np.array_equal(e, eo), T


def test_harness_matches_oracle(checker, port):
    rng = np.random.default_rng(9)
    a = rng.standard_normal(800)
    b = a + rng.standard_normal(800) * (rng.random(800) < 0.3) * 1e-3
    mean_abs, mean_rel, sig = host.computeDifferenceMetrics(a, b)
    sig_o, rel_o = checker.diff_metrics_c(a, b)
    assert (sig, mean_rel, mean_abs) == (sig_o, rel_o, 0.0)
    assert host.calculate_flops(83869696, 2.5e-4) == port.calculate_flops(83869696, 2.5e-4)
    host.initialize_metrics()
    for t in (3.0, 1.0, 2.0, 6.0):
        host.update_medium_metric(host.B200_CSR_TIME, t)
    assert host.get_metric_value(host.B200_CSR_TIME) == 3.0
    assert host.get_metric_min(host.B200_CSR_TIME) == 1.0 and host.get_metric_median(host.B200_CSR_TIME) == 2.5
    host.accumulateErrors(0.5, 0.25, host.B200_CSR_TIME)
    ma, mr = host.computeAverageErrors(host.B200_CSR_TIME)
    assert (ma, mr) == (0.5 / (4 + host.ITERATION_SKIP), 0.25 / (4 + host.ITERATION_SKIP))
    host.reset_medium_time_metrics()
    assert host.get_metric_value(host.B200_CSR_TIME) == 0.0
    host.cleanup_metrics()
    # algorithmic bytes of BASELINE.md section 3: lap2d 4096^2 -> 1.342 GB
    M, nnz = 4096 * 4096, 83869696
    assert host.calculate_csr_bytes(M, M, nnz) == nnz * 12 + 4 * (M + 1) + 16 * M
    assert abs(host.calculate_csr_bytes(M, M, nnz) / 1e9 - 1.342) < 1e-3
    v = np.zeros(7)
    assert (host.init_vector_at_one(v) == 1.0).all()


def test_driver_fails_loudly_without_a_gpu(tmp_path):
    """tools/spmv_driver.c (the main()-style driver) has no CPU fallback: on a machine without a CUDA device it must
    stop with exit status 3 and say why; on a GPU box the same command is covered by tests/test_gpu_parity.py."""
    import subprocess
    import sparsematrixvectormultiplication_b200 as pkg
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("CUDA device present")
    except ImportError:
        pass
    exe = pkg.LIB_PATH.parent / "spmv_driver"
    if not exe.exists():
        pkg.build()
    out = subprocess.run([str(exe), "--csv", str(tmp_path / "r.csv"), str(GOLDEN / "mtx" / "general_matrix.mtx")],
                         capture_output=True, text=True)
    assert out.returncode == 3, out.stdout + out.stderr
    assert "no CUDA device" in out.stderr and "no CPU fallback" in out.stderr
    assert not (tmp_path / "r.csv").exists()
