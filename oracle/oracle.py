"""ctypes bindings for the CPU ORACLE -- TEST INFRASTRUCTURE, not the product.

Two back-ends with one Python face:

* ``Restated``  -> oracle/liboracle.so   (oracle.c: plain-C restatement, every function cites the
  reference file:line it follows)
* ``Reference`` -> oracle/_ref/libspmv_ref.so  (the UNMODIFIED reference CPU sources compiled by
  oracle/Makefile from /root/reference; the .so travels to the GPU box, the sources do not)

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product package never does (tests/test_no_oracle_in_product.py
enforces it).
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import subprocess
import sys
from dataclasses import dataclass
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
HACK_SIZE = 32

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)
c_ll_p = C.POINTER(C.c_longlong)


class OracleError(RuntimeError):
    pass


def build(quiet: bool = True) -> None:
    """Compile liboracle.so and (when /root/reference is mounted) _ref/libspmv_ref.so."""
    out = subprocess.run(["make", "-C", str(HERE)], capture_output=True, text=True)
    if out.returncode != 0:
        raise OracleError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


@contextlib.contextmanager
def quiet_stdout():
    """The reference printf()s from inside its partitioners; silence fd 1 around such calls."""
    libc = C.CDLL(None)
    sys.stdout.flush()
    libc.fflush(None)
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(devnull, 1)
        yield
    finally:
        libc.fflush(None)
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


def _dp(a):
    return a.ctypes.data_as(c_dbl_p)


@dataclass
class Coo:
    M: int
    N: int
    nz: int
    I: np.ndarray
    J: np.ndarray
    val: np.ndarray
    type: str = "MCRG"


@dataclass
class Hll:
    """HLL as one flat row-major arena: slot (b, r, j) -> offset[b] + r*maxnz[b] + j."""
    num_blocks: int
    rows: np.ndarray
    maxnz: np.ndarray
    offset: np.ndarray
    JA: np.ndarray
    AS: np.ndarray

    def block(self, b):
        lo, hi = int(self.offset[b]), int(self.offset[b + 1])
        shape = (int(self.rows[b]), int(self.maxnz[b]))
        return self.JA[lo:hi].reshape(shape), self.AS[lo:hi].reshape(shape)


# ---------------------------------------------------------------------------------------------
# restated oracle (oracle.c)
# ---------------------------------------------------------------------------------------------
class _OrcCoo(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("nz", C.c_int), ("I", c_int_p), ("J", c_int_p),
                ("val", c_dbl_p), ("type", C.c_char * 4)]


class _OrcHll(C.Structure):
    _fields_ = [("num_blocks", C.c_int), ("rows", c_int_p), ("maxnz", c_int_p),
                ("offset", c_ll_p), ("JA", c_int_p), ("AS", c_dbl_p)]


class Restated:
    kind = "port"

    def __init__(self):
        path = HERE / "liboracle.so"
        if not path.exists():
            build()
        self.lib = lib = C.CDLL(str(path))
        lib.orc_read_matrix_market.argtypes = [C.c_char_p, C.POINTER(_OrcCoo)]
        lib.orc_coo_to_csr.argtypes = [C.c_int, C.c_int, c_int_p, c_int_p, c_dbl_p, c_int_p, c_int_p, c_dbl_p]
        lib.orc_coo_to_hll.argtypes = [C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_dbl_p, C.POINTER(_OrcHll)]
        lib.orc_spmv_csr_serial.argtypes = [C.c_int, c_int_p, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p]
        lib.orc_spmv_csr_serial.restype = None
        lib.orc_spmv_hll_serial.argtypes = [C.POINTER(_OrcHll), c_dbl_p, c_dbl_p]
        lib.orc_spmv_hll_serial.restype = None
        lib.orc_partition_rows.argtypes = [C.c_int, c_int_p, C.c_int, C.c_longlong, c_int_p, c_int_p]
        lib.orc_partition_hll_blocks.argtypes = [C.POINTER(_OrcHll), C.c_int, C.c_int, c_int_p, c_int_p]
        lib.orc_spmv_csr_parallel.argtypes = [c_int_p, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p, C.c_int, c_int_p, c_int_p]
        lib.orc_spmv_csr_parallel.restype = None
        lib.orc_spmv_hll_parallel.argtypes = [C.POINTER(_OrcHll), c_dbl_p, c_dbl_p, C.c_int, c_int_p, c_int_p]
        lib.orc_spmv_hll_parallel.restype = None
        lib.orc_calculate_flops.argtypes = [C.c_int, C.c_double]
        lib.orc_calculate_flops.restype = C.c_double
        lib.orc_diff_metrics_c.argtypes = [c_dbl_p, c_dbl_p, C.c_int, C.c_double, C.c_double, c_dbl_p]
        lib.orc_diff_metrics_cuda.argtypes = [c_dbl_p, c_dbl_p, C.c_int, c_dbl_p, c_dbl_p]
        lib.orc_diff_metrics_cuda.restype = None
        lib.orc_power_iteration.argtypes = [C.c_int, c_int_p, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p, C.c_int, c_dbl_p]
        lib.orc_power_iteration.restype = None

    # -- parser ------------------------------------------------------------------------------
    def read_matrix_market(self, path) -> Coo:
        raw = _OrcCoo()
        if self.lib.orc_read_matrix_market(os.fsencode(str(path)), C.byref(raw)) != 0:
            raise OracleError(f"read_matrix_market failed: {path}")
        n = raw.nz
        coo = Coo(raw.M, raw.N, n,
                  np.ctypeslib.as_array(raw.I, (n,)).copy() if n else np.zeros(0, np.int32),
                  np.ctypeslib.as_array(raw.J, (n,)).copy() if n else np.zeros(0, np.int32),
                  np.ctypeslib.as_array(raw.val, (n,)).copy() if n else np.zeros(0, np.float64),
                  raw.type.decode("ascii"))
        self.lib.orc_free_coo(C.byref(raw))
        return coo

    # -- builders ----------------------------------------------------------------------------
    def coo_to_csr(self, coo: Coo):
        I, J, V = _i32(coo.I), _i32(coo.J), _f64(coo.val)
        row_ptr = np.zeros(coo.M + 1, np.int32)
        col_idx = np.zeros(coo.nz, np.int32)
        values = np.zeros(coo.nz, np.float64)
        if self.lib.orc_coo_to_csr(coo.M, coo.nz, _ip(I), _ip(J), _dp(V), _ip(row_ptr), _ip(col_idx), _dp(values)):
            raise OracleError("coo_to_csr failed")
        return row_ptr, col_idx, values

    def coo_to_hll(self, coo: Coo) -> Hll:
        I, J, V = _i32(coo.I), _i32(coo.J), _f64(coo.val)
        raw = _OrcHll()
        if self.lib.orc_coo_to_hll(coo.M, coo.N, coo.nz, _ip(I), _ip(J), _dp(V), C.byref(raw)):
            raise OracleError("coo_to_hll failed")
        nb = raw.num_blocks
        offset = np.ctypeslib.as_array(raw.offset, (nb + 1,)).astype(np.int64)
        total = int(offset[nb])
        h = Hll(nb,
                np.ctypeslib.as_array(raw.rows, (nb,)).copy() if nb else np.zeros(0, np.int32),
                np.ctypeslib.as_array(raw.maxnz, (nb,)).copy() if nb else np.zeros(0, np.int32),
                offset,
                np.ctypeslib.as_array(raw.JA, (total,)).copy() if total else np.zeros(0, np.int32),
                np.ctypeslib.as_array(raw.AS, (total,)).copy() if total else np.zeros(0, np.float64))
        self.lib.orc_free_hll(C.byref(raw))
        return h

    def _raw_hll(self, h: Hll):
        keep = (_i32(h.rows), _i32(h.maxnz), np.ascontiguousarray(h.offset, np.int64), _i32(h.JA), _f64(h.AS))
        raw = _OrcHll(h.num_blocks, _ip(keep[0]), _ip(keep[1]), keep[2].ctypes.data_as(c_ll_p), _ip(keep[3]), _dp(keep[4]))
        return raw, keep

    # -- products ----------------------------------------------------------------------------
    def spmv_csr_serial(self, row_ptr, col_idx, values, x, y=None):
        row_ptr, col_idx, values, x = _i32(row_ptr), _i32(col_idx), _f64(values), _f64(x)
        M = len(row_ptr) - 1
        if y is None:
            y = np.zeros(M, np.float64)  # the reference accumulates: the caller zeroes y
        self.lib.orc_spmv_csr_serial(M, _ip(row_ptr), _ip(col_idx), _dp(values), _dp(x), _dp(y))
        return y

    def spmv_hll_serial(self, h: Hll, x, M=None):
        raw, keep = self._raw_hll(h)
        x = _f64(x)
        y = np.zeros(h.num_blocks * HACK_SIZE, np.float64)
        self.lib.orc_spmv_hll_serial(C.byref(raw), _dp(x), _dp(y))
        return y[: (M if M is not None else int(h.rows.sum()))]

    def partition_rows(self, row_ptr, num_threads, total_nnz=None):
        row_ptr = _i32(row_ptr)
        M = len(row_ptr) - 1
        if total_nnz is None:
            total_nnz = int(row_ptr[-1])
        T = max(int(num_threads), 1)
        start, end = np.zeros(T, np.int32), np.zeros(T, np.int32)
        used = self.lib.orc_partition_rows(M, _ip(row_ptr), num_threads, total_nnz, _ip(start), _ip(end))
        return start[:used].copy(), end[:used].copy()

    def partition_hll(self, h: Hll, N, num_threads):
        raw, keep = self._raw_hll(h)
        T = max(int(num_threads), 1)
        start, end = np.zeros(T, np.int32), np.zeros(T, np.int32)
        used = self.lib.orc_partition_hll_blocks(C.byref(raw), N, num_threads, _ip(start), _ip(end))
        return start[:used].copy(), end[:used].copy()

    def spmv_csr_parallel(self, row_ptr, col_idx, values, x, start, end, y=None):
        row_ptr, col_idx, values, x = _i32(row_ptr), _i32(col_idx), _f64(values), _f64(x)
        start, end = _i32(start), _i32(end)
        if y is None:
            y = np.zeros(len(row_ptr) - 1, np.float64)
        self.lib.orc_spmv_csr_parallel(_ip(row_ptr), _ip(col_idx), _dp(values), _dp(x), _dp(y), len(start), _ip(start), _ip(end))
        return y

    def spmv_hll_parallel(self, h: Hll, x, start, end, M=None):
        raw, keep = self._raw_hll(h)
        x, start, end = _f64(x), _i32(start), _i32(end)
        y = np.zeros(h.num_blocks * HACK_SIZE, np.float64)
        self.lib.orc_spmv_hll_parallel(C.byref(raw), _dp(x), _dp(y), len(start), _ip(start), _ip(end))
        return y[: (M if M is not None else int(h.rows.sum()))]

    # -- harness -----------------------------------------------------------------------------
    def calculate_flops(self, nz, seconds):
        return self.lib.orc_calculate_flops(nz, seconds)

    def diff_metrics_c(self, ref, res, abs_tol=1e-5, rel_tol=1e-4):
        ref, res = _f64(ref), _f64(res)
        rel = C.c_double()
        sig = self.lib.orc_diff_metrics_c(_dp(ref), _dp(res), len(ref), abs_tol, rel_tol, C.byref(rel))
        return sig, rel.value

    def diff_metrics_cuda(self, ref, res):
        ref, res = _f64(ref), _f64(res)
        a, r = C.c_double(), C.c_double()
        self.lib.orc_diff_metrics_cuda(_dp(ref), _dp(res), len(ref), C.byref(a), C.byref(r))
        return a.value, r.value

    def power_iteration(self, row_ptr, col_idx, values, x0, iters):
        row_ptr, col_idx, values = _i32(row_ptr), _i32(col_idx), _f64(values)
        x = _f64(x0).copy()
        M = len(row_ptr) - 1
        y = np.zeros(M, np.float64)
        lam = np.zeros(iters, np.float64)
        self.lib.orc_power_iteration(M, _ip(row_ptr), _ip(col_idx), _dp(values), _dp(x), _dp(y), iters, _dp(lam))
        return x, y, lam


# ---------------------------------------------------------------------------------------------
# the real reference (oracle/_ref/libspmv_ref.so); struct layouts: reference
# libs/matrix_parser.h:6-14, libs/csr_matrix.h:8-16, libs/hll_matrix.h:15-27
# ---------------------------------------------------------------------------------------------
class _PreMatrix(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("nz", C.c_int), ("I", c_int_p), ("J", c_int_p),
                ("val", c_dbl_p), ("type", C.c_char * 4)]


class _CSRMatrix(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("nz", C.c_int), ("row_ptr", c_int_p),
                ("col_idx", c_int_p), ("values", c_dbl_p), ("type", C.c_char * 4)]


class _ELLPACKBlock(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("MAXNZ", C.c_int), ("JA", c_int_p), ("AS", c_dbl_p)]


class _HLLMatrix(C.Structure):
    _fields_ = [("num_blocks", C.c_int), ("blocks", C.POINTER(_ELLPACKBlock))]


class _DiffMetrics(C.Structure):
    _fields_ = [("mean_abs_err", C.c_double), ("mean_rel_err", C.c_double), ("significant_diffs", C.c_int)]


def lap3d_csr_host(n: int):
    """The n^3 7-point Laplacian (same arrays as synth.lap3d_csr), written by all host threads; bench.py's reference arm."""
    lib = Restated().lib
    lib.orc_lap3d_csr.restype = C.c_longlong
    lib.orc_lap3d_csr.argtypes = [C.c_int, c_int_p, c_int_p, c_dbl_p]
    M = n ** 3
    row_ptr = np.empty(M + 1, np.int32)
    nnz = int(lib.orc_lap3d_csr(n, _ip(row_ptr), None, None))
    if nnz < 0:
        raise OracleError("lap3d too large for int32 indices")
    col_idx, values = np.empty(nnz, np.int32), np.empty(nnz, np.float64)
    lib.orc_lap3d_csr(n, _ip(row_ptr), _ip(col_idx), _dp(values))
    return row_ptr, col_idx, values


def norm_scale(y, x, threads: int) -> float:
    """lambda = |y|_2, x = y / lambda on `threads` host threads (OpenMP region of oracle.c)."""
    lib = Restated().lib
    lib.orc_norm_scale.restype = C.c_double
    lib.orc_norm_scale.argtypes = [c_dbl_p, c_dbl_p, C.c_longlong, C.c_int]
    return float(lib.orc_norm_scale(_dp(y), _dp(x), len(y), int(threads)))


def reference_available() -> bool:
    return (HERE / "_ref" / "libspmv_ref.so").exists()


class Reference:
    kind = "reference"

    def __init__(self):
        path = HERE / "_ref" / "libspmv_ref.so"
        if not path.exists():
            build()
        if not path.exists():
            raise OracleError("oracle/_ref/libspmv_ref.so is missing and /root/reference is not mounted")
        self.lib = lib = C.CDLL(str(path))
        lib.read_matrix_market.argtypes = [C.c_char_p, C.POINTER(_PreMatrix)]
        lib.free_pre_matrix.argtypes = [C.POINTER(_PreMatrix)]
        lib.convert_in_csr.argtypes = [C.POINTER(_PreMatrix), C.POINTER(_CSRMatrix), C.c_char_p]
        lib.free_csr_matrix.argtypes = [C.POINTER(_CSRMatrix)]
        lib.convert_to_hll.argtypes = [C.POINTER(_PreMatrix), C.POINTER(_HLLMatrix)]
        lib.free_hll_matrix.argtypes = [C.POINTER(_HLLMatrix)]
        lib.csr_matrix_vector_mult.argtypes = [C.c_int, c_int_p, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p]
        lib.csr_matrix_vector_mult.restype = None
        lib.spmv_hll_serial.argtypes = [C.c_int, C.POINTER(_ELLPACKBlock), c_dbl_p, c_dbl_p]
        lib.spmv_hll_serial.restype = None
        lib.prepare_thread_distribution.argtypes = [C.c_int, c_int_p, C.c_int, C.c_longlong,
                                                    C.POINTER(c_int_p), C.POINTER(c_int_p)]
        lib.prepare_thread_distribution_hll.argtypes = [C.POINTER(_HLLMatrix), C.c_int,
                                                        C.POINTER(c_int_p), C.POINTER(c_int_p)]
        for name in ("spvm_csr_parallel", "spvm_csr_parallel_simd"):
            f = getattr(lib, name)
            f.argtypes = [c_int_p, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p, C.c_int, c_int_p, c_int_p]
            f.restype = None
        for name in ("spmv_hll", "spmv_hll_simd"):
            f = getattr(lib, name)
            f.argtypes = [C.POINTER(_ELLPACKBlock), c_dbl_p, c_dbl_p, C.c_int, c_int_p, c_int_p]
            f.restype = None
        lib.calculate_flops.argtypes = [C.c_int, C.c_double]
        lib.calculate_flops.restype = C.c_double
        lib.computeDifferenceMetrics.argtypes = [c_dbl_p, c_dbl_p, C.c_int, C.c_double, C.c_double, C.c_bool]
        lib.computeDifferenceMetrics.restype = _DiffMetrics
        self._libc = C.CDLL(None)
        self._libc.free.argtypes = [C.c_void_p]

    def _pre(self, coo: Coo):
        keep = (_i32(coo.I), _i32(coo.J), _f64(coo.val))
        pre = _PreMatrix(coo.M, coo.N, coo.nz, _ip(keep[0]), _ip(keep[1]), _dp(keep[2]), coo.type.encode("ascii")[:4])
        return pre, keep

    def read_matrix_market(self, path) -> Coo:
        pre = _PreMatrix()
        with quiet_stdout():
            rc = self.lib.read_matrix_market(os.fsencode(str(path)), C.byref(pre))
        if rc != 0:
            raise OracleError(f"read_matrix_market failed: {path}")
        n = pre.nz
        coo = Coo(pre.M, pre.N, n,
                  np.ctypeslib.as_array(pre.I, (n,)).copy() if n else np.zeros(0, np.int32),
                  np.ctypeslib.as_array(pre.J, (n,)).copy() if n else np.zeros(0, np.int32),
                  np.ctypeslib.as_array(pre.val, (n,)).copy() if n else np.zeros(0, np.float64),
                  bytes(pre.type).decode("ascii"))
        self.lib.free_pre_matrix(C.byref(pre))
        return coo

    def coo_to_csr(self, coo: Coo):
        pre, keep = self._pre(coo)
        csr = _CSRMatrix()
        with quiet_stdout():
            rc = self.lib.convert_in_csr(C.byref(pre), C.byref(csr), b"oracle")
        if rc != 0:
            raise OracleError("convert_in_csr failed")
        out = (np.ctypeslib.as_array(csr.row_ptr, (coo.M + 1,)).copy(),
               np.ctypeslib.as_array(csr.col_idx, (coo.nz,)).copy() if coo.nz else np.zeros(0, np.int32),
               np.ctypeslib.as_array(csr.values, (coo.nz,)).copy() if coo.nz else np.zeros(0, np.float64))
        self.lib.free_csr_matrix(C.byref(csr))
        return out

    def coo_to_hll(self, coo: Coo) -> Hll:
        pre, keep = self._pre(coo)
        raw = _HLLMatrix()
        with quiet_stdout():
            rc = self.lib.convert_to_hll(C.byref(pre), C.byref(raw))
        if rc != 0:
            raise OracleError("convert_to_hll failed")
        nb = raw.num_blocks
        rows = np.zeros(nb, np.int32)
        maxnz = np.zeros(nb, np.int32)
        offset = np.zeros(nb + 1, np.int64)
        for b in range(nb):
            blk = raw.blocks[b]
            rows[b], maxnz[b] = blk.M, blk.MAXNZ
            assert blk.N == coo.N
            offset[b + 1] = offset[b] + blk.M * blk.MAXNZ
        JA = np.zeros(int(offset[nb]), np.int32)
        AS = np.zeros(int(offset[nb]), np.float64)
        for b in range(nb):
            blk = raw.blocks[b]
            n = int(offset[b + 1] - offset[b])
            if n:
                JA[offset[b]:offset[b + 1]] = np.ctypeslib.as_array(blk.JA, (n,))
                AS[offset[b]:offset[b + 1]] = np.ctypeslib.as_array(blk.AS, (n,))
            else:
                assert not blk.JA and not blk.AS  # empty block: NULL arrays (src/hll_matrix.c:121-125)
        self.lib.free_hll_matrix(C.byref(raw))
        return Hll(nb, rows, maxnz, offset, JA, AS)

    def _raw_hll(self, h: Hll, N: int):
        JA, AS = _i32(h.JA), _f64(h.AS)
        blocks = (_ELLPACKBlock * max(h.num_blocks, 1))()
        for b in range(h.num_blocks):
            lo = int(h.offset[b])
            n = int(h.offset[b + 1]) - lo
            blocks[b].M, blocks[b].N, blocks[b].MAXNZ = int(h.rows[b]), N, int(h.maxnz[b])
            if n:
                blocks[b].JA = C.cast(JA.ctypes.data + 4 * lo, c_int_p)
                blocks[b].AS = C.cast(AS.ctypes.data + 8 * lo, c_dbl_p)
        raw = _HLLMatrix(h.num_blocks, C.cast(blocks, C.POINTER(_ELLPACKBlock)))
        return raw, (JA, AS, blocks)

    def spmv_csr_serial(self, row_ptr, col_idx, values, x, y=None):
        row_ptr, col_idx, values, x = _i32(row_ptr), _i32(col_idx), _f64(values), _f64(x)
        M = len(row_ptr) - 1
        if y is None:
            y = np.zeros(M, np.float64)
        self.lib.csr_matrix_vector_mult(M, _ip(row_ptr), _ip(col_idx), _dp(values), _dp(x), _dp(y))
        return y

    def spmv_hll_serial(self, h: Hll, x, M=None, N=None):
        x = _f64(x)
        raw, keep = self._raw_hll(h, N if N is not None else len(x))
        y = np.zeros(h.num_blocks * HACK_SIZE, np.float64)
        self.lib.spmv_hll_serial(raw.num_blocks, raw.blocks, _dp(x), _dp(y))
        return y[: (M if M is not None else int(h.rows.sum()))]

    def _take_ranges(self, used, ps, pe):
        start = np.ctypeslib.as_array(ps, (used,)).copy() if used else np.zeros(0, np.int32)
        end = np.ctypeslib.as_array(pe, (used,)).copy() if used else np.zeros(0, np.int32)
        if ps:
            self._libc.free(C.cast(ps, C.c_void_p))
        if pe:
            self._libc.free(C.cast(pe, C.c_void_p))
        return start, end

    def partition_rows(self, row_ptr, num_threads, total_nnz=None):
        row_ptr = _i32(row_ptr)
        if total_nnz is None:
            total_nnz = int(row_ptr[-1])
        ps, pe = c_int_p(), c_int_p()
        with quiet_stdout():
            used = self.lib.prepare_thread_distribution(len(row_ptr) - 1, _ip(row_ptr), num_threads, total_nnz,
                                                        C.byref(ps), C.byref(pe))
        return self._take_ranges(used, ps, pe)

    def partition_hll(self, h: Hll, N, num_threads):
        raw, keep = self._raw_hll(h, N)
        ps, pe = c_int_p(), c_int_p()
        with quiet_stdout():
            used = self.lib.prepare_thread_distribution_hll(C.byref(raw), num_threads, C.byref(ps), C.byref(pe))
        return self._take_ranges(used, ps, pe)

    def spmv_csr_parallel(self, row_ptr, col_idx, values, x, start, end, y=None, simd=False):
        row_ptr, col_idx, values, x = _i32(row_ptr), _i32(col_idx), _f64(values), _f64(x)
        start, end = _i32(start), _i32(end)
        if y is None:
            y = np.zeros(len(row_ptr) - 1, np.float64)
        fn = self.lib.spvm_csr_parallel_simd if simd else self.lib.spvm_csr_parallel
        fn(_ip(row_ptr), _ip(col_idx), _dp(values), _dp(x), _dp(y), len(start), _ip(start), _ip(end))
        return y

    def spmv_hll_parallel(self, h: Hll, x, start, end, M=None, N=None, simd=False):
        x, start, end = _f64(x), _i32(start), _i32(end)
        raw, keep = self._raw_hll(h, N if N is not None else len(x))
        y = np.zeros(h.num_blocks * HACK_SIZE, np.float64)
        fn = self.lib.spmv_hll_simd if simd else self.lib.spmv_hll
        fn(raw.blocks, _dp(x), _dp(y), len(start), _ip(start), _ip(end))
        return y[: (M if M is not None else int(h.rows.sum()))]

    def calculate_flops(self, nz, seconds):
        return self.lib.calculate_flops(nz, seconds)

    def diff_metrics_c(self, ref, res, abs_tol=1e-5, rel_tol=1e-4):
        ref, res = _f64(ref), _f64(res)
        with quiet_stdout():
            d = self.lib.computeDifferenceMetrics(_dp(ref), _dp(res), len(ref), abs_tol, rel_tol, False)
        return d.significant_diffs, d.mean_rel_err


def best_available():
    """The real reference when its .so exists (this container / shipped with gpurun), else the port."""
    return Reference() if reference_available() else Restated()


# ---- the reference's own GPU kernels, recompiled for sm_100a (oracle/_ref/libspmv_ref_cuda.so) -----------------
def reference_cuda_available() -> bool:
    return (HERE / "_ref" / "libspmv_ref_cuda.so").exists()


class ReferenceCuda:
    """Launches the UNMODIFIED reference kernels (cuda_src/csr_matrix_cuda.cu:122-241, cuda_src/hll_matrix.cu:346-479)
    through oracle/ref_cuda_shim.cu, with the launch shapes of the reference driver.  The "existing kernel" bar of
    bench.py and a second checker for the GPU parity tests.  Arguments are CUDA tensors (anything with data_ptr())."""
    CSR_KERNELS = ("spmv_csr_naive_kernel", "spmv_csr_warp_kernel", "spmv_csr_warp_shared_memory_kernel")
    HLL_KERNELS = ("spmv_hll_naive_kernel", "spmv_hll_warp_kernel", "spmv_hll_warp_shared_kernel_v1")

    def __init__(self):
        self.lib = C.CDLL(str(HERE / "_ref" / "libspmv_ref_cuda.so"))
        self.lib.ref_cuda_csr_spmv.restype = C.c_int
        self.lib.ref_cuda_csr_spmv.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6
        self.lib.ref_cuda_hll_upload.restype = C.c_int
        self.lib.ref_cuda_hll_upload.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        self.lib.ref_cuda_hll_spmv.restype = C.c_int
        self.lib.ref_cuda_hll_spmv.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib.ref_cuda_hll_free.restype = None
        self.lib.ref_cuda_hll_free.argtypes = [C.c_void_p]

    def csr_spmv(self, which, M, N, row_ptr, col_idx, values, x, y, stream=None):
        rc = self.lib.ref_cuda_csr_spmv(int(which), int(M), int(N), row_ptr.data_ptr(), col_idx.data_ptr(), values.data_ptr(),
                                        x.data_ptr(), y.data_ptr(), stream)
        if rc != 0:
            raise OracleError(f"reference CUDA CSR kernel {which} failed")
        return y

    def hll_upload(self, host_hll_struct_ptr, M):
        """host_hll_struct_ptr: ctypes pointer to a host HLLMatrix in the reference layout (row-major blocks)."""
        h = C.c_void_p()
        if self.lib.ref_cuda_hll_upload(host_hll_struct_ptr, int(M), C.byref(h)) != 0:
            raise OracleError("reference CUDA HLL upload failed")
        return h

    def hll_spmv(self, handle, which, x, y, stream=None):
        if self.lib.ref_cuda_hll_spmv(handle, int(which), x.data_ptr(), y.data_ptr(), stream) != 0:
            raise OracleError(f"reference CUDA HLL kernel {which} failed")
        return y

    def hll_free(self, handle):
        self.lib.ref_cuda_hll_free(handle)
