"""GPU parity tests: the CUDA path, called through the C-ABI (libspmv_b200.so), against the CPU
oracle on the same inputs, the committed golden vectors, and -- at BASELINE.json's full sizes --
size-independent properties.

Tolerance (BASELINE.json north_star): y within 1e-12 relative (fp64) of the reference's serial CSR
product.  Checked in the cancellation-proof componentwise form |dy_i| <= 1e-12 * sum_j |a_ij x_j|
and, where the data is positive, as the plain |dy_i| / |y_i|.  Index / value arrays: bit-exact.
Rows reduced by one thread (threads_per_row = 1) must be BIT-IDENTICAL to the serial oracle.
"""
import numpy as np
import pytest

from conftest import GOLDEN, golden_names, load_golden, ramp, random_coo

pytestmark = pytest.mark.gpu

TOL = 1e-12


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def abs_row_sums(checker, rp, ci, va, x):
    return checker.spmv_csr_serial(rp, ci, np.abs(va), np.abs(x))


def assert_close(y, y_ref, scale, what=""):
    err = np.abs(y - y_ref)
    bound = TOL * np.maximum(scale, np.finfo(float).tiny)
    bad = np.nonzero(err > bound)[0]
    assert bad.size == 0, f"{what}: {bad.size} rows off, first {bad[:5]}, err {err[bad[:5]]}, scale {scale[bad[:5]]}"


@pytest.fixture(scope="module")
def dev():
    import torch
    from sparsematrixvectormultiplication_b200 import device
    assert torch.cuda.is_available()
    assert device.device_count() >= 1
    torch.cuda.set_device(0)
    return device


def csr_products(dev, A, x, M):
    """y from the automatic choice (host round trip) and from every CSR kernel explicitly."""
    import torch
    out = {}
    out["auto-host"] = A.spmv_host(x)
    xd = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    for name, algo in (("stream", dev.ALGO_STREAM), ("tile", dev.ALGO_TILE), ("vector", dev.ALGO_VECTOR),
                       ("binned", dev.ALGO_BINNED), ("row", dev.ALGO_ROW)):
        yd = torch.full((max(M, 1),), float("nan"), dtype=torch.float64, device="cuda")
        A.spmv(xd, yd, algo=algo)
        out[name] = yd.cpu().numpy()[:M]
    return out


def hll_products(dev, H, x, M):
    """y from the automatic choice (host round trip) and from both HLL kernels explicitly."""
    import torch
    out = {"auto-host": H.spmv_host(x)}
    xd = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    for name, flag in (("stream", False), ("slice", True), ("rows", "rows")):
        yd = torch.full((max(M, 1),), float("nan"), dtype=torch.float64, device="cuda")
        H.spmv(xd, yd, slice_kernel=flag)
        out[name] = yd.cpu().numpy()[:M]
    return out


# ---------------------------------------------------------------------------------------------------
# golden fixtures (BASELINE config 1 among them)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names())
def test_golden_fixture_products(dev, checker, name):
    from sparsematrixvectormultiplication_b200 import host
    g = load_golden(name)
    pre = host.read_matrix_market(GOLDEN / "mtx" / f"{name}.mtx")
    csr = host.convert_in_csr(pre)
    hll = host.convert_to_hll(pre)
    A = dev.DeviceCSR.from_host(csr)
    H = dev.DeviceHLL.from_host(hll)
    for tag, x in (("ones", np.ones(pre.N)), ("ramp", ramp(pre.N))):
        y_ref = g[f"y_csr_{tag}"]
        scale = abs_row_sums(checker, g["row_ptr"], g["col_idx"], g["values"], x)
        for path, y in csr_products(dev, A, x, pre.M).items():
            assert_close(y, y_ref, scale, f"{name}/{tag}/csr-{path}")
        A.replan(threads_per_row=1)
        assert np.array_equal(bits(A.spmv_host(x)), bits(y_ref)), "one thread per row must be bit-exact"
        A.replan(threads_per_row=0)
        hp = hll_products(dev, H, x, pre.M)
        for path, y in hp.items():
            assert_close(y, y_ref, scale, f"{name}/{tag}/hll-{path}")
        if H.info().max_maxnz <= 32:  # hacks wider than 32 columns take the split "wide hack" path (tolerance only)
            assert np.array_equal(bits(hp["stream"]), bits(g[f"y_hll_{tag}"])), "stream kernel sums each row in serial order"
        assert np.array_equal(bits(hp["rows"]), bits(g[f"y_hll_{tag}"])), "lane-per-row kernel: the order of spmv_hll_serial, always"
        assert np.array_equal(bits(csr_products(dev, A, x, pre.M)["row"]), bits(y_ref)), "thread-per-row kernel: the serial loop, always"
    # device image -> host round trip reproduces the reference arrays exactly
    back = H.download()
    rows, maxnz, offset, JA, AS = back.flat()
    assert np.array_equal(rows, g["hll_rows"]) and np.array_equal(maxnz, g["hll_maxnz"])
    assert np.array_equal(JA, g["hll_JA"]) and np.array_equal(bits(AS), bits(g["hll_AS"]))
    for b in range(back.num_blocks):
        assert (back.c.blocks[b].MAXNZ == 0) == (not back.c.blocks[b].JA)
    i = H.info()
    assert i.num_hacks == len(g["hll_rows"]) and i.nnz_reference_slots == int(g["hll_offset"][-1])


def test_config1_fixture_literal(dev):
    """expected y for general_matrix.mtx with x = 1."""
    from sparsematrixvectormultiplication_b200 import host
    pre = host.read_matrix_market(GOLDEN / "mtx" / "general_matrix.mtx")
    csr = host.convert_in_csr(pre)
    y = np.zeros(pre.M)
    host.csr_matrix_vector_mult(csr.M, csr.row_ptr, csr.col_idx, csr.values, np.ones(pre.N), y)
    expect = np.array([0, 0.49154282666738891, 0, -0.66141388497577847, 0, 0, 0, 0, 0, 0.6136755441817987])
    assert np.array_equal(y, expect)
    hll = host.convert_to_hll(pre)
    yh = np.full(hll.num_blocks * 32, np.nan)
    host.spmv_hll_serial(hll, np.ones(pre.N), yh)
    assert np.array_equal(yh[:pre.M], expect)


# ---------------------------------------------------------------------------------------------------
# random matrices: duplicates, empty rows, ragged shapes
# ---------------------------------------------------------------------------------------------------
SHAPES = [(1, 1, 1), (1, 50, 40), (50, 1, 30), (31, 33, 200), (32, 32, 0), (33, 64, 500), (64, 64, 64),
          (257, 100, 3000), (1000, 1000, 20000), (5000, 300, 9000), (777, 5000, 60000)]


@pytest.mark.parametrize("M,N,nz", SHAPES)
@pytest.mark.parametrize("dup", [False, True])
def test_random_matrices(dev, checker, M, N, nz, dup):
    from sparsematrixvectormultiplication_b200 import host
    rng = np.random.default_rng(M * 7919 + N * 31 + nz + dup)
    coo = random_coo(rng, M, N, nz, dup=dup)
    pre = host.PreMatrix(M, N, coo.I, coo.J, coo.val)
    csr = host.convert_in_csr(pre)
    hll = host.convert_to_hll(pre)
    x = rng.standard_normal(N)
    rp, ci, va = checker.coo_to_csr(coo)
    assert np.array_equal(csr.col_idx, ci) and np.array_equal(bits(csr.values), bits(va))
    y_ref = checker.spmv_csr_serial(rp, ci, va, x)
    scale = abs_row_sums(checker, rp, ci, va, x)
    A = dev.DeviceCSR.from_host(csr)
    for path, y in csr_products(dev, A, x, M).items():
        assert_close(y, y_ref, scale, f"csr-{path}")
    for tile_items, long_threshold in ((64, 8), (128, 100), (4096, 2048)):
        for tpr in (0, 1, 4, 32):
            A.replan(tile_items=tile_items, long_threshold=long_threshold, threads_per_row=tpr)
            y = A.spmv_host(x)
            assert_close(y, y_ref, scale, f"csr-tiles D={tile_items} L={long_threshold} tpr={tpr}")
            if tpr == 1 and A.info().num_long_rows == 0:
                assert np.array_equal(bits(y), bits(y_ref))
    H = dev.DeviceHLL.from_host(hll)
    y_hll_ref = checker.spmv_hll_serial(checker.coo_to_hll(coo), x, M)
    hp = hll_products(dev, H, x, M)
    for path, y in hp.items():
        assert_close(y, y_ref, scale, f"hll-{path}")
        assert_close(y, y_hll_ref, scale, f"hll-{path}-vs-hll-serial")
    if H.info().max_maxnz <= 32:  # no "wide hack" path: every row is summed in the serial order
        assert np.array_equal(bits(hp["stream"]), bits(y_hll_ref))
    if not dup:  # device-side CSR -> HLL equals convert_to_hll for duplicate-free rows
        H2 = A.to_hll()
        rows, maxnz, offset, JA, AS = H2.download().flat()
        h = checker.coo_to_hll(coo)
        assert np.array_equal(maxnz, h.maxnz) and np.array_equal(JA, h.JA) and np.array_equal(bits(AS), bits(h.AS))
        assert_close(H2.spmv_host(x), y_ref, scale, "hll-from-csr")
        assert H2.info().slots == H.info().slots and H2.info().nnz_reference_slots == H.info().nnz_reference_slots


def test_skewed_rows_take_the_long_row_path(dev, checker):
    """One row far above the long-row threshold (split over several CTAs), a few around it, many short."""
    rng = np.random.default_rng(42)
    N = 200_000
    lengths = np.concatenate([[100_000, 1025, 1024, 9000, 0, 8192, 8193], rng.integers(0, 12, 3000)])
    M = len(lengths)
    rp = np.zeros(M + 1, np.int32)
    np.cumsum(lengths, out=rp[1:])
    ci = np.concatenate([np.sort(rng.choice(N, n, replace=False)) for n in lengths]).astype(np.int32)
    va = rng.uniform(0.1, 1.0, rp[-1])
    x = rng.uniform(0.1, 1.0, N)
    y_ref = checker.spmv_csr_serial(rp, ci, va, x)
    A = dev.DeviceCSR.upload(M, N, rp, ci, va)
    A.replan(long_threshold=1024)
    info = A.info()
    assert info.num_long_rows == 5 and info.num_fragments == 13 + 1 + 2 + 1 + 2
    for path, y in csr_products(dev, A, x, M).items():
        assert np.max(np.abs(y - y_ref) / y_ref.clip(1e-300)) <= TOL, path
    y1, y2 = A.spmv_host(x), A.spmv_host(x)
    assert np.array_equal(bits(y1), bits(y2)), "long-row combine must be run-to-run deterministic"
    yacc = rng.standard_normal(M)
    expect = checker.spmv_csr_serial(rp, ci, va, x, y=yacc.copy())
    got = A.spmv_host(x, y=yacc.copy(), accumulate=True)
    assert np.max(np.abs(got - expect)) <= TOL * np.max(np.abs(expect))


def test_accumulate_is_the_reference_serial_semantics(dev, checker):
    """y += A x with one thread per row reproduces csr_matrix_vector_mult bit for bit."""
    rng = np.random.default_rng(8)
    coo = random_coo(rng, 900, 700, 12000, dup=True)
    rp, ci, va = checker.coo_to_csr(coo)
    x, y0 = rng.standard_normal(700), rng.standard_normal(900)
    expect = checker.spmv_csr_serial(rp, ci, va, x, y=y0.copy())
    A = dev.DeviceCSR.upload(900, 700, rp, ci, va).replan(threads_per_row=1)
    assert np.array_equal(bits(A.spmv_host(x, y=y0.copy(), accumulate=True)), bits(expect))


@pytest.mark.parametrize("vec", [0, 1, 2, 4, 8, 16, 32])
def test_plan_free_vector_kernel_on_raw_device_arrays(dev, checker, vec):
    import torch
    rng = np.random.default_rng(vec)
    coo = random_coo(rng, 2000, 1500, 40000, dup=False)
    rp, ci, va = checker.coo_to_csr(coo)
    x = rng.standard_normal(1500)
    t = [torch.from_numpy(a).cuda() for a in (rp, ci, va, x)]
    y = torch.empty(2000, dtype=torch.float64, device="cuda")
    dev.csr_spmv_raw(2000, *t, y, threads_per_row=vec)
    assert_close(y.cpu().numpy(), checker.spmv_csr_serial(rp, ci, va, x), abs_row_sums(checker, rp, ci, va, x))
    A = dev.DeviceCSR.wrap(2000, 1500, t[0], t[1], t[2])
    y2 = torch.zeros(2000, dtype=torch.float64, device="cuda")
    A.spmv(t[3], y2)
    assert_close(y2.cpu().numpy(), checker.spmv_csr_serial(rp, ci, va, x), abs_row_sums(checker, rp, ci, va, x))
    y3 = torch.full((2000,), -7.0, dtype=torch.float64, device="cuda")
    A.spmv_rows(100, 1234, t[3], y3)
    got = y3.cpu().numpy()
    assert (got[:100] == -7.0).all() and (got[1234:] == -7.0).all()
    assert_close(got[100:1234], checker.spmv_csr_serial(rp, ci, va, x)[100:1234],
                 abs_row_sums(checker, rp, ci, va, x)[100:1234])


# ---------------------------------------------------------------------------------------------------
# drop-in host entry points (reference signatures, GPU inside)
# ---------------------------------------------------------------------------------------------------
def test_drop_in_entry_points(dev, checker):
    from sparsematrixvectormultiplication_b200 import host
    rng = np.random.default_rng(77)
    M, N = 1234, 999
    coo = random_coo(rng, M, N, 30000, dup=True)
    pre = host.PreMatrix(M, N, coo.I, coo.J, coo.val)
    csr, hll = host.convert_in_csr(pre), host.convert_to_hll(pre)
    x = rng.standard_normal(N)
    rp, ci, va = checker.coo_to_csr(coo)
    y_ref = checker.spmv_csr_serial(rp, ci, va, x)
    scale = abs_row_sums(checker, rp, ci, va, x)
    y0 = rng.standard_normal(M)
    y = y0.copy()
    host.csr_matrix_vector_mult(M, csr.row_ptr, csr.col_idx, csr.values, x, y)  # accumulates
    assert_close(y, y0 + y_ref, scale + np.abs(y0), "csr_matrix_vector_mult")
    for T in (1, 3, 8):
        s, e = host.prepare_thread_distribution(M, csr.row_ptr, T, csr.nz)
        for fn in (host.spvm_csr_parallel, host.spvm_csr_parallel_simd):
            y = np.full(M, np.nan)
            fn(csr.row_ptr, csr.col_idx, csr.values, x, y, len(s), s, e)
            assert_close(y, y_ref, scale, fn.__name__)
        bs, be = host.prepare_thread_distribution_hll(hll, T)
        for fn in (host.spmv_hll, host.spmv_hll_simd):
            y = np.full(hll.num_blocks * 32, np.nan)
            fn(hll, x, y, len(bs), bs, be)
            assert_close(y[:M], y_ref, scale, fn.__name__)
    y = np.full(hll.num_blocks * 32, np.nan)
    host.spmv_hll_serial(hll, x, y)
    assert_close(y[:M], y_ref, scale, "spmv_hll_serial")
    # a range that covers only part of the rows leaves the others untouched (reference semantics)
    y = np.full(M, -3.0)
    host.spvm_csr_parallel(csr.row_ptr, csr.col_idx, csr.values, x, y, 1, np.array([100], np.int32), np.array([200], np.int32))
    assert (y[:100] == -3.0).all() and (y[200:] == -3.0).all()
    assert_close(y[100:200], y_ref[100:200], scale[100:200], "partial range")


# ---------------------------------------------------------------------------------------------------
# device generators == numpy twins (bit-exact)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,p0,p1,p2", [("lap2d", 37, 0, 0), ("lap2d", 1, 0, 0), ("lap3d", 11, 0, 0), ("lap3d", 2, 0, 0),
                                           ("uniform", 1000, 4096, 32), ("uniform", 333, 70, 7)])
def test_device_generators_match_numpy_twins(dev, checker, kind, p0, p1, p2):
    from sparsematrixvectormultiplication_b200 import synth
    if kind == "lap2d":
        k, twin, M = synth.SYNTH_LAP2D, lambda lo, hi: synth.lap2d_csr(p0, lo, hi), p0 * p0
    elif kind == "lap3d":
        k, twin, M = synth.SYNTH_LAP3D, lambda lo, hi: synth.lap3d_csr(p0, lo, hi), p0 ** 3
    else:
        k, twin, M = synth.SYNTH_UNIFORM, lambda lo, hi: synth.uniform_csr(p0, p1, p2, synth.DEFAULT_SEED, lo, hi), p0
    for lo, hi in ((0, M), (M // 3, M - M // 5), (M // 2, M // 2)):
        A = dev.DeviceCSR.synth(k, p0, p1, p2, seed=synth.DEFAULT_SEED, row_begin=lo, row_end=hi)
        rp, ci, va = A.download()
        trp, tci, tva = twin(lo, hi)
        assert np.array_equal(rp, trp) and np.array_equal(ci, tci) and np.array_equal(bits(va), bits(tva))
        if hi > lo:
            N = A.info().N
            x = synth.hash_vector(N, seed=99)
            assert_close(A.spmv_host(x), checker.spmv_csr_serial(trp, tci, tva, x), abs_row_sums(checker, trp, tci, tva, x))


def test_device_vector_generator_and_helpers(dev):
    import torch
    from sparsematrixvectormultiplication_b200 import synth
    n = 1_000_003
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    dev.synth_vector(x, 123)
    xh = synth.hash_vector(n, 123)
    assert np.array_equal(bits(x.cpu().numpy()), bits(xh))
    ws = torch.empty(dev.vec_ws_doubles(), dtype=torch.float64, device="cuda")
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    dev.vec_sumsq(x, ws, out)
    ss = float(out.item())
    assert abs(ss - float(np.dot(xh, xh))) <= 1e-13 * ss
    out2 = torch.zeros(1, dtype=torch.float64, device="cuda")
    dev.vec_sumsq(x, ws, out2)
    assert out2.item() == ss  # deterministic
    z = torch.empty_like(x)
    dev.vec_scale_by_inv_norm(z, x, out)
    assert np.array_equal(bits(z.cpu().numpy()), bits(xh / np.sqrt(ss)))
    dev.vec_fill(z, 1.0)
    assert (z == 1.0).all().item()


def test_power_iteration_matches_the_oracle(dev, port):
    """BASELINE config 5 semantics on a small 3-D Laplacian: y = A x; lambda = |y|; x = y / lambda."""
    import torch
    from sparsematrixvectormultiplication_b200 import synth
    n, iters = 12, 25
    rp, ci, va = synth.lap3d_csr(n)
    x_ref, y_ref, lam_ref = port.power_iteration(rp, ci, va, np.ones(n ** 3), iters)
    A = dev.DeviceCSR.synth(synth.SYNTH_LAP3D, n)
    x = torch.ones(n ** 3, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    ws = torch.empty(dev.vec_ws_doubles(), dtype=torch.float64, device="cuda")
    ss = torch.zeros(1, dtype=torch.float64, device="cuda")
    lam = []
    for _ in range(iters):
        A.spmv(x, y)
        dev.vec_sumsq(y, ws, ss)
        dev.vec_scale_by_inv_norm(x, y, ss)
        lam.append(float(ss.item()) ** 0.5)
    assert np.max(np.abs(np.array(lam) - lam_ref) / lam_ref) <= TOL
    assert np.max(np.abs(x.cpu().numpy() - x_ref)) <= TOL * np.max(np.abs(x_ref))


# ---------------------------------------------------------------------------------------------------
# BASELINE.json full sizes: oracle where it finishes in seconds, otherwise properties
# ---------------------------------------------------------------------------------------------------
def test_full_size_lap2d_4096(dev, checker):
    """Config 2: 16.8 M rows, 83.9 M nnz.  x = 1 gives exact small integers; the ramp vector is checked
    against the serial oracle on the downloaded arrays (about a second of CPU work)."""
    import torch
    from sparsematrixvectormultiplication_b200 import synth
    n = 4096
    A = dev.DeviceCSR.synth(synth.SYNTH_LAP2D, n)
    info = A.info()
    assert (info.M, info.nnz) == (n * n, 83_869_696) and info.algorithmic_bytes == 83_869_696 * 12 + 4 * (n * n + 1) + 16 * n * n
    rp, ci, va = A.download()
    assert rp[-1] == info.nnz
    x = torch.ones(n * n, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    A.spmv(x, y)
    assert np.array_equal(y.cpu().numpy(), 5.0 - np.diff(rp))  # 4 - (#neighbours), exact
    xr = ramp(n * n)
    y_ref = checker.spmv_csr_serial(rp, ci, va, xr)
    got = A.spmv_host(xr)
    assert np.max(np.abs(got - y_ref)) <= TOL * 16.0  # |A||x| <= 8 * 1.75
    A.replan(threads_per_row=1)
    assert np.array_equal(bits(A.spmv_host(xr)), bits(y_ref)), "one thread per row: bit-exact with the serial loop"
    A.replan(threads_per_row=0)
    assert np.array_equal(bits(A.spmv_host(xr)), bits(y_ref)), "5-point rows take the in-order path of the stream kernel anyway"
    A.replan(threads_per_row=1)
    H = A.to_hll()
    hi = H.info()
    # MAXNZ is 5 everywhere except the hacks of the first and last grid row (4): SURVEY.md section 8(d)
    assert hi.slots == 5 * n * n - 2 * n and hi.max_maxnz == 5 and hi.num_hacks == n * n // 32
    xd = torch.from_numpy(xr).cuda()
    ys = torch.empty_like(x)
    H.spmv(xd, ys, slice_kernel=False)
    assert np.array_equal(bits(ys.cpu().numpy()), bits(y_ref)), "HLL stream kernel sums every row in serial order (padding adds 0.0)"
    assert np.array_equal(bits(H.spmv_host(xr)), bits(y_ref)), "narrow hacks: the automatic choice is the stream kernel"
    H.spmv(xd, ys, slice_kernel=True)
    assert np.max(np.abs(ys.cpu().numpy() - y_ref)) <= TOL * 16.0
    A.spmv(xd, ys, algo=dev.ALGO_TILE)
    assert np.array_equal(bits(ys.cpu().numpy()), bits(y_ref))
    yv = torch.empty_like(x)
    A.spmv(torch.from_numpy(xr).cuda(), yv, algo=dev.ALGO_VECTOR)
    assert np.max(np.abs(yv.cpu().numpy() - y_ref)) <= TOL * 16.0


def test_full_size_uniform_8m_properties(dev, checker):
    """Config 3: 8.4 M x 8.4 M, 32 nnz/row (268 M nnz, 3.2 GB).  EVERY kernel (automatic choice, vector, binned, tile,
    stream; HLL automatic, slice, stream) against the serial oracle (reference src/csr_matrix.c:130-139) on the WHOLE
    matrix, componentwise |dy_i| <= 1e-12 * sum_j |a_ij x_j| (all terms are positive, so that sum is y_ref_i itself);
    three row windows of the downloaded arrays are also checked against the numpy twin of the generator."""
    import torch
    from sparsematrixvectormultiplication_b200 import synth
    M = N = 1 << 23
    A = dev.DeviceCSR.synth(synth.SYNTH_UNIFORM, M, N, 32)
    H = A.to_hll()
    assert A.info().nnz == 268_435_456 and H.info().slots == 268_435_456 and H.info().max_maxnz == 32
    x = torch.empty(N, dtype=torch.float64, device="cuda")
    dev.synth_vector(x, 4242)
    y_csr, y_hll, y_vec = (torch.empty(M, dtype=torch.float64, device="cuda") for _ in range(3))
    A.spmv(x, y_csr)
    H.spmv(x, y_hll)
    A.spmv(x, y_vec, algo=dev.ALGO_VECTOR)
    y_alt = torch.empty_like(y_csr)
    for flag in (True, False):
        H.spmv(x, y_alt, slice_kernel=flag)
        assert float(((y_hll - y_alt).abs() / y_hll).max()) <= TOL
    for algo in (dev.ALGO_TILE, dev.ALGO_STREAM):
        A.spmv(x, y_alt, algo=algo)
        assert float(((y_csr - y_alt).abs() / y_csr).max()) <= TOL
    # all values are positive: plain relative error is well posed
    assert float(((y_csr - y_hll).abs() / y_csr).max()) <= TOL
    assert float(((y_csr - y_vec).abs() / y_csr).max()) <= TOL
    xh = synth.hash_vector(N, 4242)
    assert np.array_equal(x.cpu().numpy(), xh)
    rp_all, ci_all, va_all = A.download()
    for lo in (0, 1_234_567, M - 4096):                          # the device generator == its numpy twin
        rp, ci, va = synth.uniform_csr(M, N, 32, synth.DEFAULT_SEED, lo, lo + 4096)
        assert np.array_equal(ci_all[32 * lo: 32 * (lo + 4096)], ci) and np.array_equal(va_all[32 * lo: 32 * (lo + 4096)], va)
        assert np.array_equal(rp_all[lo: lo + 4097] - rp_all[lo], rp)
    y_ref = checker.spmv_csr_serial(rp_all, ci_all, va_all, xh)   # the whole matrix through the serial oracle
    del rp_all, ci_all, va_all
    assert y_ref.min() > 0
    y_ref_d = torch.from_numpy(y_ref).cuda()

    def whole(y_dev, what):
        worst = float(((y_dev - y_ref_d).abs() / y_ref_d).max())
        assert worst <= TOL, f"{what}: max componentwise error {worst:.3e} vs the serial oracle on all {M} rows"

    whole(y_csr, "CSR automatic choice")
    whole(y_hll, "HLL automatic choice")
    whole(y_vec, "CSR vector kernel")
    for algo in (dev.ALGO_BINNED, dev.ALGO_TILE, dev.ALGO_STREAM):
        A.spmv(x, y_alt, algo=algo)
        whole(y_alt, f"CSR algo {algo}")
    for flag in (True, False):
        H.spmv(x, y_alt, slice_kernel=flag)
        whole(y_alt, f"HLL slice_kernel={flag}")
    import os
    os.environ["SPMV_B200_L2_PERSIST"] = "1"                      # the persisting-L2 window changes no bit
    try:
        A.spmv(x, y_alt, algo=dev.ALGO_VECTOR)
        assert torch.equal(y_alt, y_vec)
        H.spmv(x, y_alt, slice_kernel=True)
        whole(y_alt, "HLL slice kernel with the L2 window")
    finally:
        os.environ.pop("SPMV_B200_L2_PERSIST", None)
    # linearity: A(2x) = 2 A x exactly (power-of-two scaling commutes with rounding)
    x2 = x * 2.0
    y2 = torch.empty_like(y_csr)
    A.spmv(x2, y2)
    assert torch.equal(y2, y_csr * 2.0)


def test_full_size_rmat_24_properties(dev, checker):
    """Config 4: R-MAT scale 24, 16 edges per row on average, longest row > 100 000 (skewed-row load balancing).
    The automatic choice (row-binned kernel, long-row fragments in the same launch) and the tile / stream kernels against
    the serial oracle (reference src/csr_matrix.c:130-139) on the WHOLE matrix, componentwise
    |dy_i| <= 1e-12 * sum_j |a_ij x_j| (all terms positive: that sum is y_ref_i); an independent torch index_add
    reference, linearity and run-to-run determinism on top."""
    import torch
    from sparsematrixvectormultiplication_b200 import synth
    scale = 24
    M = 1 << scale
    rp, ci, va = synth.rmat_csr_device(scale, 16)
    lengths = rp[1:] - rp[:-1]
    assert int(lengths.max()) > 100_000 and int(rp[-1]) > 200_000_000
    A = dev.DeviceCSR.wrap(M, M, rp, ci, va)
    info = A.info()
    assert info.auto_algo == dev.ALGO_BINNED and info.max_row_nnz == int(lengths.max())
    x = torch.empty(M, dtype=torch.float64, device="cuda")
    dev.synth_vector(x, 777)                                    # x in (0, 1], values in (0, 1]: everything positive
    y = torch.empty(M, dtype=torch.float64, device="cuda")
    A.spmv(x, y)
    rows = torch.repeat_interleave(torch.arange(M, device="cuda"), lengths.long())
    ref = torch.zeros(M, dtype=torch.float64, device="cuda").index_add_(0, rows, va * x[ci.long()])
    del rows
    nonempty = ref > 0
    assert float(((y - ref).abs()[nonempty] / ref[nonempty]).max()) <= 4 * TOL, "binned kernel vs torch index_add reference"
    assert bool((y[~nonempty] == 0).all()), "empty rows give exactly 0"
    y2 = torch.empty_like(y)
    A.spmv(x, y2)
    assert torch.equal(y, y2), "run-to-run deterministic (fixed-order long-row combine, no atomics)"
    for algo in (dev.ALGO_TILE, dev.ALGO_STREAM):
        A.spmv(x, y2, algo=algo)
        assert float(((y - y2).abs()[nonempty] / ref[nonempty]).max()) <= 4 * TOL, f"algo {algo}"
    A.spmv(x * 2.0, y2)
    assert torch.equal(y2, y * 2.0), "linearity under power-of-two scaling is exact"
    # the whole matrix through the serial oracle
    y_ref = checker.spmv_csr_serial(rp.cpu().numpy(), ci.cpu().numpy(), va.cpu().numpy(), x.cpu().numpy())
    y_ref_d = torch.from_numpy(y_ref).cuda()
    assert bool(((y_ref_d > 0) == nonempty).all())
    for algo, what in ((dev.ALGO_AUTO, "automatic choice (binned)"), (dev.ALGO_TILE, "tile"), (dev.ALGO_STREAM, "stream")):
        A.spmv(x, y2, algo=algo)
        worst = float(((y2 - y_ref_d).abs()[nonempty] / y_ref_d[nonempty]).max())
        assert worst <= TOL, f"{what}: max componentwise error {worst:.3e} vs the serial oracle on all {M} rows"
        assert bool((y2[~nonempty] == 0).all())
    longest = int(lengths.argmax())
    assert abs(float(y[longest]) - y_ref[longest]) <= TOL * y_ref[longest], f"longest row ({int(lengths.max())} nonzeros)"
    A.close()


def test_full_size_lap3d_512(dev, checker):
    """Config 5's matrix on ONE GPU: 134 M rows, 938 M nnz (11.8 GB).  x = 1 gives exact small integers (6 - number of
    neighbours); with the ramp vector the automatic choice, the row kernel and the stream kernel must agree bit for bit
    (all sum rows of 7 in the serial order), HLL too; a window of rows is checked against the serial oracle."""
    import torch
    from sparsematrixvectormultiplication_b200 import synth
    n = 512
    A = dev.DeviceCSR.synth(synth.SYNTH_LAP3D, n)
    info = A.info()
    assert (info.M, info.nnz) == (n ** 3, 937_951_232) and info.max_row_nnz == 7
    x = torch.ones(n ** 3, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    A.spmv(x, y)
    # 6 on the diagonal, -1 per neighbour: the row sum is the number of missing neighbours (0..3), exact
    idx = torch.arange(n ** 3, device="cuda")
    i, j, k = idx // (n * n), (idx // n) % n, idx % n
    missing = sum(((c == 0).long() + (c == n - 1).long()) for c in (i, j, k)).double()
    assert torch.equal(y, missing)
    del idx, i, j, k, missing
    xr = 1.0 + (torch.arange(n ** 3, device="cuda") % 7).double() / 8.0
    y_auto, y_alt = torch.empty_like(x), torch.empty_like(x)
    A.spmv(xr, y_auto)
    for algo in (dev.ALGO_ROW, dev.ALGO_STREAM):
        A.spmv(xr, y_alt, algo=algo)
        assert torch.equal(y_auto, y_alt), f"algo {algo}: rows of 7 are summed in the serial order by every stencil kernel"
    lo = 77 * n * n + 5
    rp, ci, va = synth.lap3d_csr(n, lo, lo + 5000)            # numpy twin of the device generator, rows [lo, lo + 5000)
    y_ref = checker.spmv_csr_serial(rp, ci, va, xr.cpu().numpy())
    assert np.array_equal(bits(y_auto[lo:lo + 5000].cpu().numpy()), bits(y_ref)), "bit-exact with csr_matrix_vector_mult"
    H = A.to_hll()
    assert H.info().max_maxnz == 7
    H.spmv(xr, y_alt)
    assert torch.equal(y_auto, y_alt), "HLL (padding adds 0.0) gives the same bits"
    H.close()
    A.close()


def test_fused_power_iteration_matches_the_oracle(dev, port):
    """One launch per iteration (product + lazy normalisation + |w|^2 partials) against the oracle's
    y = A x; lambda = |y|; x = y / lambda."""
    import torch
    from sparsematrixvectormultiplication_b200 import synth
    from sparsematrixvectormultiplication_b200.distributed import FusedPowerIteration
    n, iters = 14, 30
    rp, ci, va = synth.lap3d_csr(n)
    x_ref, y_ref, lam_ref = port.power_iteration(rp, ci, va, np.ones(n ** 3), iters)
    F = FusedPowerIteration(synth.SYNTH_LAP3D, n)
    lam = []
    for _ in range(iters):
        F.step()
        lam.append(F.eigenvalue_estimate())
    assert np.max(np.abs(np.array(lam) - lam_ref) / lam_ref) <= TOL
    v = F.normalized_x().cpu().numpy()
    assert np.max(np.abs(v - x_ref)) <= TOL * np.max(np.abs(x_ref))
    # the fused entry point with no scaling and no partials is the plain product
    A = dev.DeviceCSR.synth(synth.SYNTH_LAP3D, n)
    x = torch.from_numpy(ramp(n ** 3)).cuda()
    y1, y2 = torch.empty_like(x), torch.empty_like(x)
    A.spmv(x, y1)
    A.spmv_fused(x, y2)
    assert torch.equal(y1, y2)


@pytest.mark.parametrize("mailbox", [False, True])
def test_hll_fused_power_iteration_is_bitwise_the_csr_iteration(dev, port, mailbox, monkeypatch):
    """The HLL twin of the fused iterated product (hll_row_fused_kernel): against the oracle's power iteration within
    1e-12, and BITWISE against the CSR iteration (same thread -> row mapping, same partial sums; padding slots add +0.0).
    n = 15: 3375 rows = 105 full hacks + one short hack of 15 rows, rows of 4..7 nonzeros, hacks of width 6 and 7."""
    import torch
    from sparsematrixvectormultiplication_b200 import synth
    from sparsematrixvectormultiplication_b200.distributed import FusedPowerIteration
    n, iters = 15, 25
    rp, ci, va = synth.lap3d_csr(n)
    x_ref, _, lam_ref = port.power_iteration(rp, ci, va, np.ones(n ** 3), iters)
    monkeypatch.setenv("SPMV_B200_FUSED_BATCH", "4")   # the CSR side: fused ROW kernel (small matrices default to the fused stream kernel, whose partials are grouped differently)
    Fh = FusedPowerIteration(synth.SYNTH_LAP3D, n, fmt="hll", mailbox=mailbox)
    Fc = FusedPowerIteration(synth.SYNTH_LAP3D, n, fmt="csr", mailbox=mailbox)
    for _ in range(iters):
        Fh.step()
        Fc.step()
        assert Fh.eigenvalue_estimate() == Fc.eigenvalue_estimate()
    assert abs(Fh.eigenvalue_estimate() - lam_ref[-1]) <= TOL * lam_ref[-1]
    vh, vc = Fh.normalized_x(), Fc.normalized_x()
    assert torch.equal(vh, vc)
    assert np.max(np.abs(vh.cpu().numpy() - x_ref)) <= TOL * np.max(np.abs(x_ref))
    # no scaling, no partials: the plain HLL product in the order of spmv_hll_serial
    A = dev.DeviceCSR.synth(synth.SYNTH_LAP3D, n)
    H = A.to_hll()
    x = torch.from_numpy(ramp(n ** 3)).cuda()
    y1, y2 = torch.empty_like(x), torch.full_like(x, float("nan"))
    H.spmv(x, y1, slice_kernel="rows")
    H.spmv_fused(x, y2)
    assert torch.equal(y1, y2)
    # the partials buffer may be larger than the grid of this launch: the tail is zeroed by the launch itself
    partials = torch.full((H.partials_count() + 7,), float("nan"), dtype=torch.float64, device="cuda")
    H.spmv_fused(x, y2, partials=partials[:H.partials_count()])
    assert bool(torch.isfinite(partials[:H.partials_count()]).all())
    assert abs(float(partials[:H.partials_count()].sum()) - float((y1 * y1).sum())) <= 1e-12 * float((y1 * y1).sum())
    pc = torch.full((A.partials_count(),), float("nan"), dtype=torch.float64, device="cuda")
    A.spmv_fused(x, y2, partials=pc)
    assert bool(torch.isfinite(pc).all()), "entries the fused launch does not own must be zeroed, not left as they were"
    assert abs(float(pc.sum()) - float((y1 * y1).sum())) <= 1e-12 * float((y1 * y1).sum())


@pytest.mark.parametrize("fmt", ["csr", "hll"])
def test_two_launch_iterated_product_matches_the_oracle(dev, port, fmt):
    """The two-launch form (FLAT fused product + one-CTA exchange kernel) against the oracle's power iteration and against
    the one-launch mailbox form: x bitwise (same rows, same order), lambda within 1e-12 (the partials are grouped per CTA,
    and the CTAs differ)."""
    import torch
    from sparsematrixvectormultiplication_b200 import synth
    from sparsematrixvectormultiplication_b200.distributed import FusedPowerIteration
    n, iters = 15, 25
    rp, ci, va = synth.lap3d_csr(n)
    x_ref, _, lam_ref = port.power_iteration(rp, ci, va, np.ones(n ** 3), iters)
    S = FusedPowerIteration(synth.SYNTH_LAP3D, n, fmt=fmt, split=True)
    F = FusedPowerIteration(synth.SYNTH_LAP3D, n, fmt=fmt, mailbox=True)
    lam = []
    for _ in range(iters):
        S.step()
        F.step()
        lam.append(S.eigenvalue_estimate())
    assert np.max(np.abs(np.array(lam) - lam_ref) / lam_ref) <= TOL
    assert abs(S.eigenvalue_estimate() - F.eigenvalue_estimate()) <= TOL * lam_ref[-1]
    vs = S.normalized_x().cpu().numpy()
    assert np.max(np.abs(vs - x_ref)) <= TOL * np.max(np.abs(x_ref))
    assert np.max(np.abs(vs - F.normalized_x().cpu().numpy())) <= TOL * np.max(np.abs(x_ref))
    S.reset(1.0)                                          # and again after a reset
    for _ in range(3):
        S.step()
    assert abs(S.eigenvalue_estimate() - lam_ref[2]) <= TOL * lam_ref[2]
    info = S.A.info()
    assert info.flat_chunks >= 1 and info.flat_batch >= 2


def test_padded_allgather_layout_remap_and_interior_rows(dev, checker):
    """The one-collective allgather refresh keeps x in a padded rank-major layout and rewrites the column indices once
    (spmv_b200_csr_remap_columns); the product on the remapped matrix with the padded x gives the bits of the original
    product, and interior_rows finds the rows that only touch the rank's own slice."""
    import torch
    from sparsematrixvectormultiplication_b200 import partition, synth
    n, world = 12, 3
    parts = partition.synth_partition(synth.SYNTH_LAP3D, n, 0, 0, world)
    starts = [s for s, _ in parts] + [n ** 3]
    stride = (max(e - s for s, e in parts) + 31) // 32 * 32
    xfull = ramp(n ** 3)
    xpad = np.zeros(world * stride)
    for p, (s, e) in enumerate(parts):
        xpad[p * stride: p * stride + e - s] = xfull[s:e]
    for rank, (lo, hi) in enumerate(parts):
        rp, ci, va = synth.lap3d_csr(n, lo, hi)
        y_ref = checker.spmv_csr_serial(rp, ci, va, xfull)
        A = dev.DeviceCSR.synth(synth.SYNTH_LAP3D, n, row_begin=lo, row_end=hi)
        ilo, ihi = A.interior_rows(lo, hi)
        inside = np.array([ci[rp[r]] >= lo and ci[rp[r + 1] - 1] < hi for r in range(hi - lo)])
        assert inside[ilo:ihi].all() and (ilo == 0 or not inside[ilo - 1]) and (ihi == hi - lo or not inside[ihi])
        assert ihi - ilo >= (hi - lo) - 2 * n * n              # all but one plane on each side
        A.remap_columns(starts, stride)
        assert A.info().N == world * stride
        _, ci2, _ = A.download()
        owner = np.searchsorted(np.array(starts), ci, side="right") - 1
        assert np.array_equal(ci2, owner * stride + ci - np.array(starts)[owner])
        y = torch.full((hi - lo,), float("nan"), dtype=torch.float64, device="cuda")
        A.spmv(torch.from_numpy(xpad).cuda(), y)
        assert np.array_equal(bits(y.cpu().numpy()), bits(y_ref))
        y.fill_(float("nan"))
        A.spmv_rows(ilo, ihi, torch.from_numpy(xpad).cuda(), y)
        got = y.cpu().numpy()
        assert np.array_equal(bits(got[ilo:ihi]), bits(y_ref[ilo:ihi])) and np.isnan(got[:ilo]).all() and np.isnan(got[ihi:]).all()


# ---------------------------------------------------------------------------------------------------
# the pipelined host entry point: any number of row windows gives the bits of the resident product
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("windows", [1, 3, 16])
@pytest.mark.parametrize("kind", ["banded", "random", "skewed"])
def test_host_pipeline_windows(dev, checker, monkeypatch, windows, kind):
    """spmv_b200_{csr,hll}_spmv_host cut the rows into windows and overlap upload / product / download
    (csrc/cuda/hostpath.cu).  The result must not depend on the number of windows, for banded matrices (windows
    start before all of x has landed), random columns (every window waits for the last chunk) and plans with
    long rows (single window)."""
    import torch
    monkeypatch.setenv("SPMV_B200_HOST_WINDOWS", str(windows))
    rng = np.random.default_rng(windows * 10 + len(kind))
    M = N = 6000
    if kind == "banded":
        offs = np.array([-70, -1, 0, 1, 70])
        rows = np.repeat(np.arange(M), len(offs))
        cols = rows + np.tile(offs, M)
        keep = (cols >= 0) & (cols < N)
        rows, cols = rows[keep], cols[keep]
    elif kind == "random":
        rows = np.repeat(np.arange(M), 20)
        cols = rng.integers(0, N, rows.size)
    else:
        lengths = np.concatenate([[5000, 3000], rng.integers(0, 9, M - 2)])
        rows = np.repeat(np.arange(M), lengths)
        cols = np.concatenate([np.sort(rng.choice(N, n, replace=False)) for n in lengths])
    order = np.lexsort((cols, rows))
    rows, cols = rows[order], cols[order]
    rp = np.zeros(M + 1, np.int32)
    np.cumsum(np.bincount(rows, minlength=M), out=rp[1:])
    ci = cols.astype(np.int32)
    va = rng.uniform(0.5, 1.5, ci.size)
    x = rng.uniform(0.5, 1.5, N)
    y_ref = checker.spmv_csr_serial(rp, ci, va, x)
    scale = abs_row_sums(checker, rp, ci, va, x)
    xd = torch.from_numpy(x).cuda()
    for algo in (dev.ALGO_AUTO, dev.ALGO_STREAM, dev.ALGO_TILE, dev.ALGO_VECTOR, dev.ALGO_BINNED, dev.ALGO_ROW):
        A = dev.DeviceCSR.upload(M, N, rp, ci, va)   # fresh handle: the window plan is made on the first host call
        yd = torch.empty(M, dtype=torch.float64, device="cuda")
        A.spmv(xd, yd, algo=algo)
        resident = yd.cpu().numpy()
        assert_close(resident, y_ref, scale, f"{kind}/algo{algo}")
        got = A.spmv_host(x, algo=algo)
        assert np.array_equal(bits(got), bits(resident)), f"{kind}/algo{algo}/W{windows}: host pipeline differs from the resident product"
        y0 = rng.standard_normal(M)
        A.spmv(xd, yd.copy_(torch.from_numpy(y0)), accumulate=True, algo=algo)
        got = A.spmv_host(x, y=y0.copy(), accumulate=True, algo=algo)
        assert np.array_equal(bits(got), bits(yd.cpu().numpy())), f"{kind}/algo{algo}/W{windows}: accumulate"
        if kind != "skewed":
            H = A.to_hll()
            for flag in (False, True):
                H.spmv(xd, yd, slice_kernel=flag)
            got = H.spmv_host(x)
            auto = torch.empty(M, dtype=torch.float64, device="cuda")
            H.spmv(xd, auto)
            assert np.array_equal(bits(got), bits(auto.cpu().numpy())), f"{kind}/hll/W{windows}"
            H.close()
        A.close()


@pytest.mark.parametrize("kind", ["lap2d", "uniform8"])
def test_host_pipeline_tapered_windows(dev, checker, kind):
    """Vectors of 16+ units of 2 MiB take the tapered window schedule of hostpath.cu (1, 1, 2, 4, 8, 16 ..., 8, 4, 2, 1, 1
    units, upload pieces ending at each window's last referenced column).  Banded columns (lap2d: windows start as soon
    as their piece lands) and random columns (every window needs all of x): the host call gives the bits of the resident
    product and matches the serial oracle."""
    import os
    import torch
    from sparsematrixvectormultiplication_b200 import synth
    if kind == "lap2d":
        n = 2100                                     # 4.41 M rows = 17 units
        A = dev.DeviceCSR.synth(synth.SYNTH_LAP2D, n)
    else:
        M = 17 * 262144 + 777
        A = dev.DeviceCSR.synth(synth.SYNTH_UNIFORM, M, M, 8)
    info = A.info()
    x = ramp(info.N)
    rp, ci, va = A.download()
    y_ref = checker.spmv_csr_serial(rp, ci, va, x)
    scale = abs_row_sums(checker, rp, ci, va, x)
    xd = torch.from_numpy(x).cuda()
    yd = torch.full((info.M,), float("nan"), dtype=torch.float64, device="cuda")
    A.spmv(xd, yd)
    resident = yd.cpu().numpy()
    assert_close(resident, y_ref, scale, kind)
    for _ in range(2):                               # the second call reuses the window plan
        got = A.spmv_host(x, np.full(info.M, np.nan))
        assert np.array_equal(bits(got), bits(resident)), f"{kind}: tapered host pipeline differs from the resident product"
    # pinned host buffers: the kernels store y straight into the caller's buffer (no download copies); and the same call
    # with that path switched off
    xh = torch.from_numpy(x).pin_memory()
    for knob in ("1", "0"):
        os.environ["SPMV_B200_HOST_ZEROCOPY"] = knob
        yh = torch.full((info.M,), float("nan"), dtype=torch.float64).pin_memory()
        A.spmv_host_ptr(xh.data_ptr(), yh.data_ptr())
        assert np.array_equal(bits(yh.numpy()), bits(resident)), f"{kind}: pinned host call, zero-copy {knob}"
    os.environ.pop("SPMV_B200_HOST_ZEROCOPY", None)
    H = A.to_hll()
    H.spmv(xd, yd)
    got = H.spmv_host(x, np.full(info.M, np.nan))
    assert np.array_equal(bits(got), bits(yd.cpu().numpy())), f"{kind}: HLL host pipeline"
    assert_close(got, y_ref, scale, kind + " hll")
    yh = torch.full((info.M,), float("nan"), dtype=torch.float64).pin_memory()
    H.spmv_host_ptr(xh.data_ptr(), yh.data_ptr())
    assert np.array_equal(bits(yh.numpy()), bits(yd.cpu().numpy())), f"{kind}: HLL pinned host call"


def test_main_style_driver_writes_the_csv(dev, tmp_path):
    """tools/spmv_driver.c: parser -> converters -> upload -> timed products of every kernel -> self-check against the
    serial-order product -> one CSV row per matrix (the reference driver's loop, main_cuda.cu:40-720)."""
    import csv
    import subprocess
    import sparsematrixvectormultiplication_b200 as pkg
    exe = pkg.LIB_PATH.parent / "spmv_driver"
    assert exe.exists(), "build() makes the driver next to the library"
    out_csv = tmp_path / "result_b200.csv"
    files = [str(GOLDEN / "mtx" / f"{n}.mtx") for n in ("general_matrix", "rand_symmetric_64x64", "rand_longrow_65x3000")]
    import torch
    gpus = min(2, torch.cuda.device_count())
    run = subprocess.run([str(exe), "--csv", str(out_csv), "--iters", "7", "--warmup", "2", "--gpus", str(gpus), "--lap2d", "300", *files],
                         capture_output=True, text=True)
    assert run.returncode == 0, run.stdout[-3000:] + run.stderr[-3000:]
    rows = list(csv.DictReader(open(out_csv)))
    assert [r["matrix_name"] for r in rows] == ["general_matrix.mtx", "rand_symmetric_64x64.mtx", "rand_longrow_65x3000.mtx", "lap2d_300"]
    assert rows[0]["rows"] == "10" and rows[0]["nonzeros"] == "5"
    assert rows[3]["rows"] == "90000" and rows[3]["nonzeros"] == str(5 * 90000 - 4 * 300)
    for r in rows:
        for k in ("csr_auto", "csr_row", "csr_stream", "csr_vector", "csr_binned", "hll_auto", "hll_rows", "hll_stream", "hll_slice"):
            assert float(r[f"time_{k}"]) > 0 and float(r[f"flops_{k}"]) > 0
            assert float(r[f"relative_error_{k}"]) <= 1e-12 and float(r[f"absolute_error_{k}"]) <= 1e-9
        assert float(r["time_e2e_csr_host"]) > 0 and r["check_baseline"] == "gpu_serial_order_kernel"
    # --gpus N: the square matrices from files also went through spmv_b200_multi_* (product + 23 power iterations)
    sym = rows[1]
    assert 1 <= int(sym["ngpus"]) <= gpus and float(sym["time_multi_iteration"]) > 0 and float(sym["relative_error_multi_product"]) <= 1e-12
    from oracle import oracle as O
    port = O.Restated()
    rp, ci, va = port.coo_to_csr(port.read_matrix_market(GOLDEN / "mtx" / "rand_symmetric_64x64.mtx"))
    _, _, lam_ref = port.power_iteration(rp, ci, va, np.ones(64), 23)
    assert abs(float(sym["lambda_multi"]) - lam_ref[-1]) <= 1e-11 * abs(lam_ref[-1])
    assert rows[2]["ngpus"] == "1" and float(rows[2]["time_multi_iteration"]) == 0.0      # 65 x 3000: not square, skipped
    assert float(rows[3]["time_multi_iteration"]) == 0.0                                   # generated on the device: no host arrays


def test_short_rows_are_summed_in_serial_order_on_every_path(dev, checker):
    """Rows of up to 12 nonzeros are reduced by one lane, left to right, with mul and add rounded separately -- the
    reference's serial loop (src/csr_matrix.c:134-138) bit for bit -- in the stream and tile kernels whatever tile,
    path (per-chunk or CTA-wide two-phase) or row partition they fall into; the binned kernel does so for its
    one-lane class (<= 8 nonzeros).  Longer rows: tolerance."""
    import torch
    rng = np.random.default_rng(2024)
    M, N = 9000, 7000
    lengths = rng.choice([0, 1, 3, 5, 6, 7, 12, 13, 40, 200, 700], size=M, p=[.05, .1, .2, .2, .1, .1, .1, .05, .05, .03, .02])
    rp = np.zeros(M + 1, np.int32)
    np.cumsum(lengths, out=rp[1:])
    ci = np.concatenate([np.sort(rng.choice(N, n, replace=False)) for n in lengths]).astype(np.int32)
    va = rng.standard_normal(rp[-1])
    x = rng.standard_normal(N)
    y_ref = checker.spmv_csr_serial(rp, ci, va, x)
    scale = abs_row_sums(checker, rp, ci, va, x)
    short, tiny = lengths <= 12, lengths <= 8
    xd = torch.from_numpy(x).cuda()
    for lo, hi in ((0, M), (1234, 7777), (8990, M)):          # the whole matrix and two row blocks of it ("ranks")
        sub_rp = (rp[lo:hi + 1] - rp[lo]).astype(np.int32)
        A = dev.DeviceCSR.upload(hi - lo, N, sub_rp, ci[rp[lo]:rp[hi]], va[rp[lo]:rp[hi]])
        for D, L in ((0, 0), (96, 16), (600, 64), (3000, 512)):
            A.replan(tile_items=D, long_threshold=L)
            for name, algo in (("stream", dev.ALGO_STREAM), ("tile", dev.ALGO_TILE), ("binned", dev.ALGO_BINNED), ("auto", dev.ALGO_AUTO),
                               ("row", dev.ALGO_ROW)):
                yd = torch.full((hi - lo,), float("nan"), dtype=torch.float64, device="cuda")
                A.spmv(xd, yd, algo=algo)
                y = yd.cpu().numpy()
                assert_close(y, y_ref[lo:hi], scale[lo:hi], f"{name} rows[{lo},{hi}) D={D} L={L}")
                exact = tiny[lo:hi] if name in ("binned", "auto") else (np.ones(hi - lo, bool) if name == "row" else short[lo:hi])
                assert np.array_equal(bits(y[exact]), bits(y_ref[lo:hi][exact])), f"{name} rows[{lo},{hi}) D={D} L={L}: short rows not bit-exact"
        A.close()


def test_reference_cuda_kernels_agree(dev, checker):
    """The reference's own GPU kernels (cuda_src/csr_matrix_cuda.cu:122-241, cuda_src/hll_matrix.cu:346-479), compiled
    unmodified for sm_100a into oracle/_ref/libspmv_ref_cuda.so, as a second checker: same matrix, same x, every new
    kernel within 1e-12 of every reference kernel's result (both are within 1e-12 of the serial CPU product).
    Row counts are multiples of 96 on purpose: spmv_csr_warp_shared_memory_kernel returns from out-of-range warps BEFORE
    its __syncthreads() and before they have loaded their share of the x cache (cuda_src/csr_matrix_cuda.cu:207-217), so
    with a ragged last block its last rows read uninitialised shared memory (seen here: 20 rows off by O(1) at M = 3000).
    That is the reference's defect, not a parity target."""
    import ctypes as C
    import torch
    from oracle import oracle as O
    from sparsematrixvectormultiplication_b200 import host
    if not O.reference_cuda_available():
        pytest.skip("oracle/_ref/libspmv_ref_cuda.so not built (needs /root/reference at build time)")
    ref = O.ReferenceCuda()
    rng = np.random.default_rng(77)
    for M, N, nz in ((3072, 2500, 60000), (288, 4000, 30000)):
        coo = random_coo(rng, M, N, nz, dup=False)
        pre = host.PreMatrix(M, N, coo.I, coo.J, coo.val)
        csr = host.convert_in_csr(pre)
        hll = host.convert_to_hll(pre)
        x = rng.standard_normal(N)
        y_cpu = checker.spmv_csr_serial(csr.row_ptr, csr.col_idx, csr.values, x)
        scale = abs_row_sums(checker, csr.row_ptr, csr.col_idx, csr.values, x)
        t = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (csr.row_ptr, csr.col_idx, csr.values, x)]
        A = dev.DeviceCSR.from_host(csr)
        H = dev.DeviceHLL.from_host(hll)
        ours = csr_products(dev, A, x, M)
        ours.update({f"hll-{k}": v for k, v in hll_products(dev, H, x, M).items()})
        theirs = {}
        for which, name in enumerate(O.ReferenceCuda.CSR_KERNELS):
            y = torch.full((M,), float("nan"), dtype=torch.float64, device="cuda")
            ref.csr_spmv(which, M, N, *t, y)
            theirs[name] = y.cpu().numpy()
        handle = ref.hll_upload(C.byref(hll.c), M)
        for which, name in enumerate(O.ReferenceCuda.HLL_KERNELS):
            y = torch.full((M,), float("nan"), dtype=torch.float64, device="cuda")
            ref.hll_spmv(handle, which, t[3], y)
            theirs[name] = y.cpu().numpy()
        ref.hll_free(handle)
        for rname, yr in theirs.items():
            assert_close(yr, y_cpu, scale, f"reference {rname} vs serial CPU")
            for oname, yo in ours.items():
                err = np.abs(yo - yr)
                assert np.all(err <= 2 * TOL * np.maximum(scale, np.finfo(float).tiny)), f"{oname} vs reference {rname}"


@pytest.mark.parametrize("M,N,nz", [(1, 1, 1), (40, 1, 25), (1, 300, 120), (33, 65, 700), (1000, 1000, 30000), (4000, 123, 50000),
                                    (64, 64, 0)])
@pytest.mark.parametrize("dup", [False, True])
def test_device_coo_to_csr(dev, checker, M, N, nz, dup):
    """spmv_b200_csr_from_coo: duplicate-free COO gives convert_in_csr's arrays bit for bit (and then CSR -> HLL on the
    device gives convert_to_hll's); with repeated coordinates the rows hold the same (column, value) multiset, sorted
    by column, and the product agrees within tolerance.  Host arrays and device tensors."""
    import torch
    from sparsematrixvectormultiplication_b200 import host
    rng = np.random.default_rng(M * 13 + N * 7 + nz + dup)
    coo = random_coo(rng, M, N, nz, dup=dup)
    nz = len(coo.I)
    rp, ci, va = checker.coo_to_csr(coo)
    x = rng.standard_normal(N)
    y_ref = checker.spmv_csr_serial(rp, ci, va, x)
    scale = abs_row_sums(checker, rp, ci, va, x)
    built = [dev.DeviceCSR.from_coo(M, N, coo.I, coo.J, coo.val)]
    if nz:
        t = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (coo.I, coo.J, coo.val)]
        built.append(dev.DeviceCSR.from_coo(M, N, *t))
    for A in built:
        drp, dci, dva = A.download()
        assert np.array_equal(drp, rp)
        if not dup:
            assert np.array_equal(dci, ci) and np.array_equal(bits(dva), bits(va))
            h = checker.coo_to_hll(coo)
            rows, maxnz, offset, JA, AS = A.to_hll().download().flat()
            assert np.array_equal(maxnz, h.maxnz) and np.array_equal(JA, h.JA) and np.array_equal(bits(AS), bits(h.AS))
        else:
            assert np.array_equal(dci, ci), "columns are sorted inside every row in both"
            for r in range(M):   # same multiset of (column, value) per row
                a = sorted(zip(dci[drp[r]:drp[r + 1]].tolist(), dva[drp[r]:drp[r + 1]].tolist()))
                b = sorted(zip(ci[rp[r]:rp[r + 1]].tolist(), va[rp[r]:rp[r + 1]].tolist()))
                assert a == b, f"row {r}"
        assert_close(A.spmv_host(x), y_ref, scale, "product of the device-built CSR")
    with pytest.raises(Exception):
        dev.DeviceCSR.from_coo(3, 3, np.array([0, 3], np.int32), np.array([0, 0], np.int32), np.array([1.0, 2.0]))


# ---------------------------------------------------------------------------------------------------
# fp32 storage, fp64 arithmetic: y within 1e-5 relative of the serial fp64 product (BASELINE.json north star)
# ---------------------------------------------------------------------------------------------------
TOL32 = 1e-5


@pytest.mark.parametrize("kind", ["stencil", "random", "skewed"])
def test_fp32_path_within_1e_5(dev, checker, kind):
    """Values, x and y stored as float32, every product formed and summed in double and rounded once on the store:
    |dy_i| <= 1e-5 * sum_j |a_ij x_j| against csr_matrix_vector_mult on the ORIGINAL fp64 data, for the row kernel
    (stencil), the vector kernel (even rows) and the binned kernel with long-row fragments (skewed); CSR and HLL."""
    import torch
    rng = np.random.default_rng(len(kind))
    if kind == "stencil":
        from sparsematrixvectormultiplication_b200 import synth
        rp, ci, va = synth.lap2d_csr(70)
        M = N = 4900
        va = va * rng.uniform(0.5, 1.5, va.size)
    else:
        M, N = 6000, 5000
        lengths = rng.integers(20, 40, M) if kind == "random" else np.concatenate([[9000, 3000, 600], rng.integers(0, 30, M - 3)])
        lengths = np.minimum(lengths, N)
        rp = np.zeros(M + 1, np.int32)
        np.cumsum(lengths, out=rp[1:])
        ci = np.concatenate([np.sort(rng.choice(N, n, replace=False)) for n in lengths]).astype(np.int32)
        va = rng.standard_normal(rp[-1])
    x = rng.standard_normal(N)
    y_ref = checker.spmv_csr_serial(rp, ci, va, x)
    scale = abs_row_sums(checker, rp, ci, va, x)
    A = dev.DeviceCSR.upload(M, N, rp, ci, va).enable_f32()
    if kind == "skewed":
        A.replan(long_threshold=512)
    x32 = torch.from_numpy(x.astype(np.float32)).cuda()
    algos = [dev.ALGO_AUTO, dev.ALGO_VECTOR, dev.ALGO_BINNED] + ([dev.ALGO_ROW] if kind != "skewed" else [])
    for algo in algos:
        y32 = torch.full((M,), float("nan"), dtype=torch.float32, device="cuda")
        A.spmv_f32(x32, y32, algo=algo)
        err = np.abs(y32.cpu().numpy().astype(np.float64) - y_ref)
        assert np.all(err <= TOL32 * np.maximum(scale, np.finfo(np.float32).tiny)), f"{kind} csr algo {algo}: max {err.max():.3g}"
        assert err.max() > 0 or kind == "stencil", "fp32 storage cannot be exact on random data"
    yh = A.spmv_host_f32(x)
    assert yh.dtype == np.float32 and np.all(np.abs(yh.astype(np.float64) - y_ref) <= TOL32 * np.maximum(scale, 1e-30))
    acc0 = rng.standard_normal(M).astype(np.float32)
    yacc = torch.from_numpy(acc0.copy()).cuda()
    A.spmv_f32(x32, yacc, accumulate=True)
    assert np.all(np.abs(yacc.cpu().numpy().astype(np.float64) - (acc0 + y_ref)) <= TOL32 * (scale + np.abs(acc0) + 1e-30))
    with pytest.raises(Exception):
        A.spmv_f32(x32, yacc, algo=dev.ALGO_STREAM)     # no fp32 stream kernel: fails loudly
    if kind != "skewed":
        H = A.to_hll().enable_f32()
        y32 = torch.full((M,), float("nan"), dtype=torch.float32, device="cuda")
        H.spmv_f32(x32, y32)
        assert np.all(np.abs(y32.cpu().numpy().astype(np.float64) - y_ref) <= TOL32 * np.maximum(scale, 1e-30)), f"{kind} hll"
        assert np.all(np.abs(H.spmv_host_f32(x).astype(np.float64) - y_ref) <= TOL32 * np.maximum(scale, 1e-30))
    B = dev.DeviceCSR.upload(M, N, rp, ci, va)
    with pytest.raises(Exception):
        B.spmv_f32(x32, yacc)                            # enable_f32 not called


def test_multi_row_forms_give_the_bits_of_the_one_row_kernels(dev, checker, monkeypatch):
    """Every (rows per thread | hacks per warp, batch, CTAs per SM) form of csr_rowm_kernel / hll_rowm_kernel, forced
    through SPMV_B200_ROW_MULTI (read at every launch): fp64 results bit for bit the reference's serial loop
    (src/csr_matrix.c:134-138) on ragged rows -- empty, 1..12 and a few longer ones, a row count that is no multiple of
    256 * rows, the whole matrix, row windows and accumulate -- and fp32 results bit for bit those of the one-row fp32
    kernel; HLL forms against the lane-per-row kernel (itself pinned to spmv_hll_serial by the tests above)."""
    import torch
    rng = np.random.default_rng(77)
    cases = []
    M, N = 9001, 7000
    lengths = rng.choice([0, 1, 3, 5, 6, 7, 12, 13, 40], size=M, p=[.08, .1, .2, .25, .1, .1, .1, .05, .02])
    lengths[-3:] = [12, 0, 7]
    cases.append((M, N, lengths))
    cases.append((1, 5, np.array([3])))
    cases.append((300, 300, np.zeros(300, np.int64)))            # no nonzero at all: x is never read
    cases.append((777, 64, rng.integers(4, 8, 777)))             # stencil-like: every form in its intended regime
    cases.append("lap2d")                                        # regular image (three runs of equal-width hacks): hll_rowu_kernel
    csr_forms, hll_forms = dev.row_forms(dev.FORMAT_CSR), dev.row_forms(dev.FORMAT_HLL)
    assert len(csr_forms) >= 8 and len(hll_forms) >= 8
    for case in cases:
        if case == "lap2d":
            from sparsematrixvectormultiplication_b200 import synth
            rp, ci, va = synth.lap2d_csr(70)
            M = N = 4900
            va = va * rng.uniform(0.5, 1.5, va.size)
        else:
            M, N, lengths = case
            rp = np.zeros(M + 1, np.int32)
            np.cumsum(lengths, out=rp[1:])
            ci = (np.concatenate([np.sort(rng.choice(N, n, replace=False)) for n in lengths]).astype(np.int32)
                  if rp[-1] else np.zeros(0, np.int32))
            va = rng.standard_normal(rp[-1])
        x = rng.standard_normal(N)
        y_ref = checker.spmv_csr_serial(rp, ci, va, x)
        acc0 = rng.standard_normal(M)
        A = dev.DeviceCSR.upload(M, N, rp, ci, va).enable_f32()
        H = A.to_hll().enable_f32()
        xd = torch.from_numpy(x).cuda()
        x32 = xd.float()
        monkeypatch.delenv("SPMV_B200_ROW_MULTI", raising=False)

        def products():
            out = {}
            y = torch.full((M,), float("nan"), dtype=torch.float64, device="cuda")
            A.spmv(xd, y, algo=dev.ALGO_ROW)
            out["csr"] = y.cpu().numpy()
            y = torch.from_numpy(acc0.copy()).cuda()
            A.spmv(xd, y, accumulate=True, algo=dev.ALGO_ROW)
            out["csr accumulate"] = y.cpu().numpy()
            y = torch.full((M,), float("nan"), dtype=torch.float64, device="cuda")
            lo, hi = M // 3, max(M // 3, M - 5)
            A.spmv_rows(lo, hi, xd, y)
            out["csr rows window"] = y.cpu().numpy()[lo:hi]
            y32 = torch.full((M,), float("nan"), dtype=torch.float32, device="cuda")
            A.spmv_f32(x32, y32, algo=dev.ALGO_ROW)
            out["csr f32"] = y32.cpu().numpy()
            y = torch.full((M,), float("nan"), dtype=torch.float64, device="cuda")
            H.spmv(xd, y, slice_kernel="rows")
            out["hll"] = y.cpu().numpy()
            if H.info().max_maxnz <= 12:
                y32 = torch.full((M,), float("nan"), dtype=torch.float32, device="cuda")
                H.spmv_f32(x32, y32)
                out["hll f32"] = y32.cpu().numpy()
            return out

        plain = products()
        assert np.array_equal(bits(plain["csr"]), bits(y_ref))
        # beyond a format's last form SPMV_B200_ROW_MULTI means its last form again; SPMV_B200_HLL_UNIFORM=B: the HLL row
        # launches of a regular image go through hll_rowu_kernel<B> (offsets by arithmetic), other images are unaffected
        settings = [{"SPMV_B200_ROW_MULTI": str(k + 1)} for k in range(max(len(csr_forms), len(hll_forms)))]
        settings += [{"SPMV_B200_HLL_UNIFORM": str(b)} for b in (3, 4, 5, 6, 7)]
        for setting in settings:
            for key, value in setting.items():
                monkeypatch.setenv(key, value)
            forced = products()
            for name, y in forced.items():
                same = (np.array_equal(y.view(np.uint32), plain[name].view(np.uint32)) if y.dtype == np.float32
                        else np.array_equal(bits(y), bits(plain[name])))
                assert same, f"{setting}, {name}, M={M}"
            assert np.array_equal(bits(forced["csr"]), bits(y_ref))
            for key in setting:
                monkeypatch.delenv(key, raising=False)
        H.close()
        A.close()


def test_resident_cache_of_the_drop_in_api(dev, checker):
    """The reference's stateless product signatures upload the matrix on every call; with the opt-in cache
    (spmv_b200_resident_cache) the second call on the same arrays reuses the device copy.  Same results, the
    cached copy dies with free_csr_matrix / free_hll_matrix, new arrays are detected."""
    import time
    from sparsematrixvectormultiplication_b200 import _native as N
    from sparsematrixvectormultiplication_b200 import host
    rng = np.random.default_rng(5)
    M = Ncols = 200_000
    coo = random_coo(rng, M, Ncols, 2_000_000, dup=False)
    pre = host.PreMatrix(M, Ncols, coo.I, coo.J, coo.val)
    csr = host.convert_in_csr(pre)
    hll = host.convert_to_hll(pre)
    x = rng.standard_normal(Ncols)
    y_ref = checker.spmv_csr_serial(csr.row_ptr, csr.col_idx, csr.values, x)
    scale = abs_row_sums(checker, csr.row_ptr, csr.col_idx, csr.values, x)
    before = N.lib().spmv_b200_resident_cache(1)
    try:
        times = []
        for _ in range(3):
            y = np.zeros(M)
            t0 = time.perf_counter()
            host.csr_matrix_vector_mult(csr.M, csr.row_ptr, csr.col_idx, csr.values, x, y)
            times.append(time.perf_counter() - t0)
            assert_close(y, y_ref, scale, "cached csr_matrix_vector_mult")
        assert min(times[1:]) < times[0], f"cached calls are not faster: {times}"
        starts, ends = host.prepare_thread_distribution(csr.M, csr.row_ptr, 8, csr.nz)
        y = np.full(M, np.nan)
        host.spvm_csr_parallel(csr.row_ptr, csr.col_idx, csr.values, x, y, len(starts), starts, ends)
        assert_close(y, y_ref, scale, "cached spvm_csr_parallel")
        yh = np.full(hll.num_blocks * 32, np.nan)
        for _ in range(2):
            host.spmv_hll_serial(hll, x, yh)
            assert_close(yh[:M], y_ref, scale, "cached spmv_hll_serial")
        bs, be = host.prepare_thread_distribution_hll(hll, 8)
        yh[:] = np.nan
        host.spmv_hll(hll, x, yh, len(bs), bs, be)
        assert_close(yh[:M], y_ref, scale, "cached spmv_hll over 8 block ranges: one upload, one product")
        # another matrix in NEW arrays is not confused with the cached one
        csr2 = host.convert_in_csr(host.PreMatrix(M, Ncols, coo.I, coo.J, -coo.val))
        y = np.zeros(M)
        host.csr_matrix_vector_mult(csr2.M, csr2.row_ptr, csr2.col_idx, csr2.values, x, y)
        assert_close(y, -y_ref, scale, "a different matrix after a cached one")
        # strict mode (2): every element is hashed, so an in-place change at a position the sampled fingerprint never
        # looks at is seen and the device copy is rebuilt
        N.lib().spmv_b200_resident_cache(2)
        y = np.zeros(M)
        host.csr_matrix_vector_mult(csr.M, csr.row_ptr, csr.col_idx, csr.values, x, y)
        assert_close(y, y_ref, scale, "strict cache, first call")
        k = 12345                                        # not a multiple of nnz // 64: unsampled
        row = int(np.searchsorted(csr.row_ptr, k, side="right") - 1)
        old = float(csr.values[k])
        csr.values[k] = old + 1000.0
        y = np.zeros(M)
        host.csr_matrix_vector_mult(csr.M, csr.row_ptr, csr.col_idx, csr.values, x, y)
        assert abs((y[row] - y_ref[row]) - 1000.0 * x[csr.col_idx[k]]) <= 1e-9 * (1.0 + abs(1000.0 * x[csr.col_idx[k]])), \
            "strict cache must notice an in-place change"
        csr.values[k] = old
    finally:
        N.lib().spmv_b200_resident_cache(before)
        N.lib().spmv_b200_resident_drop()
