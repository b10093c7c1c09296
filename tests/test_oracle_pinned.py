"""Pins the CPU oracle (oracle/oracle.c) to the reference: golden vectors produced by the unmodified
reference sources (tests/golden/make_golden.py) and, where oracle/_ref exists, live comparison with
the reference's own functions on randomised inputs."""
import json

import numpy as np
import pytest

from conftest import GOLDEN, golden_names, load_golden, ramp, random_coo

THREADS = (1, 2, 3, 4, 8, 40)


@pytest.mark.parametrize("name", golden_names())
def test_port_matches_golden(port, name):
    g = load_golden(name)
    coo = port.read_matrix_market(GOLDEN / "mtx" / f"{name}.mtx")
    assert (coo.M, coo.N, coo.nz) == (int(g["M"]), int(g["N"]), int(g["nz"]))
    assert coo.type == bytes(g["type"]).decode()
    for f in ("I", "J", "val"):
        assert np.array_equal(getattr(coo, f), g[f]), f
    rp, ci, va = port.coo_to_csr(coo)
    assert np.array_equal(rp, g["row_ptr"]) and np.array_equal(ci, g["col_idx"])
    assert np.array_equal(va.view(np.uint64), g["values"].view(np.uint64))  # bit-exact
    h = port.coo_to_hll(coo)
    for mine, gold in ((h.rows, "hll_rows"), (h.maxnz, "hll_maxnz"), (h.offset, "hll_offset"), (h.JA, "hll_JA")):
        assert np.array_equal(mine, g[gold]), gold
    assert np.array_equal(h.AS.view(np.uint64), g["hll_AS"].view(np.uint64))
    for tag, x in (("ones", np.ones(coo.N)), ("ramp", ramp(coo.N))):
        assert np.array_equal(port.spmv_csr_serial(rp, ci, va, x), g[f"y_csr_{tag}"])
        assert np.array_equal(port.spmv_hll_serial(h, x, coo.M), g[f"y_hll_{tag}"])
    for T in THREADS:
        s, e = port.partition_rows(rp, T)
        assert np.array_equal(np.stack([s, e]).reshape(2, -1), g[f"part_rows_{T}"]), T
        s, e = port.partition_hll(h, coo.N, T)
        assert np.array_equal(np.stack([s, e]).reshape(2, -1), g[f"part_hll_{T}"]), T


def test_port_rejects_what_the_reference_rejects(port):
    from oracle import oracle as O
    for name in json.loads((GOLDEN / "errors.json").read_text()):
        with pytest.raises(O.OracleError):
            port.read_matrix_market(GOLDEN / "mtx" / name)


def test_survey_golden_values(port):
    """The literal vectors quoted in SURVEY.md section 4 / BASELINE.md section 4."""
    coo = port.read_matrix_market(GOLDEN / "mtx" / "general_matrix.mtx")
    assert list(coo.I) == [1, 9, 3, 3, 9] and list(coo.J) == [0, 1, 5, 7, 9]
    rp, ci, va = port.coo_to_csr(coo)
    assert list(rp) == [0, 0, 1, 1, 3, 3, 3, 3, 3, 3, 5] and list(ci) == [0, 5, 7, 1, 9]
    y = port.spmv_csr_serial(rp, ci, va, np.ones(10))
    expect = [0, 0.49154282666738891, 0, -0.66141388497577847, 0, 0, 0, 0, 0, 0.6136755441817987]
    assert np.array_equal(y, np.array(expect))
    h = port.coo_to_hll(coo)
    assert list(h.JA) == [0, 0, 0, 0, 0, 0, 5, 7, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 9]
    s, e = port.partition_rows(rp, 4)
    assert list(s) == [0, 4] and list(e) == [4, 10]


def test_port_matches_live_reference_on_random_inputs(port, reference):
    rng = np.random.default_rng(20251018)
    for trial in range(150):
        M, N = int(rng.integers(1, 130)), int(rng.integers(1, 60))
        coo = random_coo(rng, M, N, int(rng.integers(0, 600)), dup=bool(trial % 2))
        a, b = reference.coo_to_csr(coo), port.coo_to_csr(coo)
        for u, v in zip(a, b):
            assert np.array_equal(u, v), trial
        ha, hb = reference.coo_to_hll(coo), port.coo_to_hll(coo)
        for f in ("rows", "maxnz", "offset", "JA", "AS"):
            assert np.array_equal(getattr(ha, f), getattr(hb, f)), (trial, f)
        x = rng.standard_normal(N)
        ya = reference.spmv_csr_serial(*a, x)
        assert np.array_equal(ya, port.spmv_csr_serial(*b, x))
        assert np.array_equal(reference.spmv_hll_serial(ha, x, M), port.spmv_hll_serial(hb, x, M))
        for T in (1, 2, 5, 16):
            pa, pb = reference.partition_rows(a[0], T), port.partition_rows(b[0], T)
            assert np.array_equal(pa[0], pb[0]) and np.array_equal(pa[1], pb[1]), (trial, T)
            qa, qb = reference.partition_hll(ha, N, T), port.partition_hll(hb, N, T)
            assert np.array_equal(qa[0], qb[0]) and np.array_equal(qa[1], qb[1]), (trial, T)
            if len(pa[0]):
                yp = reference.spmv_csr_parallel(*a, x, *pa)
                assert np.array_equal(yp, port.spmv_csr_parallel(*b, x, *pb))
            if len(qa[0]):
                yh = reference.spmv_hll_parallel(ha, x, *qa, M=M)
                assert np.array_equal(yh, port.spmv_hll_parallel(hb, x, *qb, M=M))


def test_port_harness_matches_reference(port, reference):
    rng = np.random.default_rng(3)
    a = rng.standard_normal(500)
    b = a + rng.standard_normal(500) * (rng.random(500) < 0.2) * 1e-3
    assert port.diff_metrics_c(a, b) == reference.diff_metrics_c(a, b)
    assert port.calculate_flops(12345, 0.5) == reference.calculate_flops(12345, 0.5)


def test_host_generator_and_norm_scale_of_the_reference_arm():
    """bench.py --impl reference at N > 1 builds the FULL 512^3 Laplacian on the host with the oracle's OpenMP generator and
    does norm + scale in an OpenMP region: both against their numpy definitions."""
    from oracle import oracle as O
    from sparsematrixvectormultiplication_b200 import synth
    for n in (1, 2, 3, 7, 20):
        got, want = O.lap3d_csr_host(n), synth.lap3d_csr(n)
        assert all(np.array_equal(a, b) and a.dtype == b.dtype for a, b in zip(got, want)), n
    rng = np.random.default_rng(3)
    y = rng.standard_normal(100_003)
    x = np.empty_like(y)
    lam = O.norm_scale(y, x, 4)
    assert abs(lam - np.linalg.norm(y)) <= 1e-13 * lam and np.array_equal(x, y / lam)
    assert O.norm_scale(y, x, 1) == lam          # fixed chunking: the sum does not depend on the thread count
