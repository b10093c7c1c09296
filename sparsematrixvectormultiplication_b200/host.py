"""Python face of the drop-in host API -- same names, argument meaning and error behaviour as the
reference's C prototypes (reference libs/matrix_parser.h:16-19, libs/csr_matrix.h:19-33,
libs/hll_matrix.h:29-39, libs/performance_calculate.h:46-67).  Everything here calls straight into
``libspmv_b200.so``; builders run on the host, products run on the GPU (no CPU fallback).

The C functions return 0 / -1 (partitioners: number of ranges, 0 on failure); the wrappers raise
``HostApiError`` on -1 so tests read like assertions on the reference's return codes.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _native as N
from ._native import HACK_SIZE  # noqa: F401

ITERATION_SKIP = 5  # reference libs/utility.h:7


class HostApiError(RuntimeError):
    pass


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ip(a):
    return a.ctypes.data_as(N.c_int_p)


def _dp(a):
    return a.ctypes.data_as(N.c_dbl_p)


def _view(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ptr, (n,))


class PreMatrix:
    """COO matrix (reference PreMatrix).  Owns C memory when produced by read_matrix_market."""

    def __init__(self, M=0, N_=0, I=None, J=None, val=None, type_="MCRG"):
        self._keep = None
        self._owned = False
        self.c = N.PreMatrixStruct()
        N.lib().init_pre_matrix(C.byref(self.c))
        if I is not None:
            I, J, val = _i32(I), _i32(J), _f64(val)
            self._keep = (I, J, val)
            self.c.M, self.c.N, self.c.nz = int(M), int(N_), len(I)
            self.c.I, self.c.J, self.c.val = _ip(I), _ip(J), _dp(val)
            self.c.type = type_.encode("ascii")[:4]

    M = property(lambda s: s.c.M)
    N = property(lambda s: s.c.N)
    nz = property(lambda s: s.c.nz)
    I = property(lambda s: _view(s.c.I, s.c.nz, np.int32))
    J = property(lambda s: _view(s.c.J, s.c.nz, np.int32))
    val = property(lambda s: _view(s.c.val, s.c.nz, np.float64))
    type = property(lambda s: bytes(s.c.type).decode("ascii"))

    def __del__(self):
        try:
            if getattr(self, "_owned", False):
                self._owned = False
                N.lib().free_pre_matrix(C.byref(self.c))
        except Exception:  # interpreter shutdown
            pass


class CSRMatrix:
    """reference CSRMatrix; arrays are numpy views of the C-owned memory."""

    def __init__(self):
        self.c = N.CSRMatrixStruct()
        N.lib().init_csr_matrix(C.byref(self.c))
        self._owned = False

    M = property(lambda s: s.c.M)
    N = property(lambda s: s.c.N)
    nz = property(lambda s: s.c.nz)
    row_ptr = property(lambda s: _view(s.c.row_ptr, s.c.M + 1, np.int32))
    col_idx = property(lambda s: _view(s.c.col_idx, s.c.nz, np.int32))
    values = property(lambda s: _view(s.c.values, s.c.nz, np.float64))

    def __del__(self):
        try:
            if getattr(self, "_owned", False):
                self._owned = False
                N.lib().free_csr_matrix(C.byref(self.c))
        except Exception:  # interpreter shutdown
            pass


class HLLMatrix:
    """reference HLLMatrix: row-major ELLPACK blocks of HACK_SIZE rows, last block short."""

    def __init__(self):
        self.c = N.HLLMatrixStruct()
        N.lib().init_hll_matrix(C.byref(self.c))
        self._owned = False
        self.rows_total = 0
        self.cols = 0

    num_blocks = property(lambda s: s.c.num_blocks)

    def block(self, b):
        blk = self.c.blocks[b]
        n = blk.M * blk.MAXNZ
        return blk.M, blk.N, blk.MAXNZ, _view(blk.JA, n, np.int32).reshape(blk.M, blk.MAXNZ), \
            _view(blk.AS, n, np.float64).reshape(blk.M, blk.MAXNZ)

    def flat(self):
        """(rows[], maxnz[], offset[], JA, AS) in one row-major arena -- the oracle's layout."""
        nb = self.num_blocks
        rows = np.array([self.c.blocks[b].M for b in range(nb)], np.int32)
        maxnz = np.array([self.c.blocks[b].MAXNZ for b in range(nb)], np.int32)
        offset = np.zeros(nb + 1, np.int64)
        np.cumsum(rows.astype(np.int64) * maxnz, out=offset[1:])
        JA = np.zeros(int(offset[nb]), np.int32)
        AS = np.zeros(int(offset[nb]), np.float64)
        for b in range(nb):
            blk = self.c.blocks[b]
            n = blk.M * blk.MAXNZ
            if n:
                JA[offset[b]:offset[b + 1]] = np.ctypeslib.as_array(blk.JA, (n,))
                AS[offset[b]:offset[b + 1]] = np.ctypeslib.as_array(blk.AS, (n,))
        return rows, maxnz, offset, JA, AS

    def __del__(self):
        try:
            if getattr(self, "_owned", False):
                self._owned = False
                N.lib().free_hll_matrix(C.byref(self.c))
        except Exception:  # interpreter shutdown
            pass


# ---- parser ---------------------------------------------------------------------------------------
def read_matrix_market(filename) -> PreMatrix:
    pre = PreMatrix()
    rc = N.lib().read_matrix_market(os.fsencode(str(filename)), C.byref(pre.c))
    if rc != 0:
        raise HostApiError(f"read_matrix_market({filename}) returned {rc}")
    pre._owned = True
    return pre


# ---- builders -------------------------------------------------------------------------------------
def convert_in_csr(pre: PreMatrix, matrix_name: str = "") -> CSRMatrix:
    csr = CSRMatrix()
    rc = N.lib().convert_in_csr(C.byref(pre.c), C.byref(csr.c), matrix_name.encode())
    if rc != 0:
        raise HostApiError(f"convert_in_csr returned {rc}")
    csr._owned = True
    return csr


def convert_to_hll(pre: PreMatrix) -> HLLMatrix:
    hll = HLLMatrix()
    rc = N.lib().convert_to_hll(C.byref(pre.c), C.byref(hll.c))
    if rc != 0:
        raise HostApiError(f"convert_to_hll returned {rc}")
    hll._owned = True
    hll.rows_total, hll.cols = pre.M, pre.N
    return hll


def _take(used, ps, pe):
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    start = np.ctypeslib.as_array(ps, (used,)).copy() if used and ps else np.zeros(0, np.int32)
    end = np.ctypeslib.as_array(pe, (used,)).copy() if used and pe else np.zeros(0, np.int32)
    if ps:
        libc.free(C.cast(ps, C.c_void_p))
    if pe:
        libc.free(C.cast(pe, C.c_void_p))
    return start, end


def prepare_thread_distribution(num_row, row_ptr, num_threads, total_nnz):
    """-> (thread_row_start, thread_row_end); their length is the C function's return value."""
    row_ptr = _i32(row_ptr)
    ps, pe = N.c_int_p(), N.c_int_p()
    used = N.lib().prepare_thread_distribution(int(num_row), _ip(row_ptr), int(num_threads), int(total_nnz),
                                               C.byref(ps), C.byref(pe))
    return _take(used, ps, pe)


def prepare_thread_distribution_hll(hll: HLLMatrix, num_threads):
    ps, pe = N.c_int_p(), N.c_int_p()
    used = N.lib().prepare_thread_distribution_hll(C.byref(hll.c), int(num_threads), C.byref(ps), C.byref(pe))
    return _take(used, ps, pe)


# ---- products (GPU through the C-ABI) ---------------------------------------------------------------
def _poisoned(y) -> bool:
    return bool(np.isnan(y).any())


def csr_matrix_vector_mult(num_row, row_ptr, col_idx, values, x, y):
    """y += A x, in place on the float64 array ``y`` (reference src/csr_matrix.c:130-139)."""
    row_ptr, col_idx, values, x = _i32(row_ptr), _i32(col_idx), _f64(values), _f64(x)
    assert y.dtype == np.float64 and y.flags.c_contiguous
    N.clear_error()      # a stale message of an earlier call must not turn a NaN in x into a failure
    N.lib().csr_matrix_vector_mult(int(num_row), _ip(row_ptr), _ip(col_idx), _dp(values), _dp(x), _dp(y))
    if num_row and _poisoned(y[:num_row]) and N.last_error():
        raise N.SpmvError(-2, N.last_error())
    return y


def _csr_ranged(fn, row_ptr, col_idx, values, x, y, starts, ends):
    row_ptr, col_idx, values, x = _i32(row_ptr), _i32(col_idx), _f64(values), _f64(x)
    starts, ends = _i32(starts), _i32(ends)
    assert y.dtype == np.float64 and y.flags.c_contiguous
    N.clear_error()
    fn(_ip(row_ptr), _ip(col_idx), _dp(values), _dp(x), _dp(y), len(starts), _ip(starts), _ip(ends))
    if len(starts) and _poisoned(y) and N.last_error():
        raise N.SpmvError(-2, N.last_error())
    return y


def spvm_csr_parallel(row_ptr, col_idx, values, x, y, num_threads, thread_row_start, thread_row_end):
    return _csr_ranged(N.lib().spvm_csr_parallel, row_ptr, col_idx, values, x, y,
                       thread_row_start[:num_threads], thread_row_end[:num_threads])


def spvm_csr_parallel_simd(row_ptr, col_idx, values, x, y, num_threads, thread_row_start, thread_row_end):
    return _csr_ranged(N.lib().spvm_csr_parallel_simd, row_ptr, col_idx, values, x, y,
                       thread_row_start[:num_threads], thread_row_end[:num_threads])


def spmv_hll_serial(hll: HLLMatrix, x, y):
    """y[32 b + i] = (A x) for every block (reference src/hll_matrix.c:286-308)."""
    x = _f64(x)
    assert y.dtype == np.float64 and y.flags.c_contiguous
    N.clear_error()
    N.lib().spmv_hll_serial(hll.c.num_blocks, hll.c.blocks, _dp(x), _dp(y))
    if _poisoned(y) and N.last_error():
        raise N.SpmvError(-2, N.last_error())
    return y


def _hll_ranged(fn, hll, x, y, starts, ends):
    x, starts, ends = _f64(x), _i32(starts), _i32(ends)
    assert y.dtype == np.float64 and y.flags.c_contiguous
    N.clear_error()
    fn(hll.c.blocks, _dp(x), _dp(y), len(starts), _ip(starts), _ip(ends))
    if len(starts) and _poisoned(y) and N.last_error():
        raise N.SpmvError(-2, N.last_error())
    return y


def spmv_hll(hll: HLLMatrix, x, y, num_threads, thread_block_start, thread_block_end):
    return _hll_ranged(N.lib().spmv_hll, hll, x, y, thread_block_start[:num_threads], thread_block_end[:num_threads])


def spmv_hll_simd(hll: HLLMatrix, x, y, num_threads, thread_block_start, thread_block_end):
    return _hll_ranged(N.lib().spmv_hll_simd, hll, x, y, thread_block_start[:num_threads],
                       thread_block_end[:num_threads])


# ---- harness ---------------------------------------------------------------------------------------
METRICS = ("SERIAL_TIME", "PARALLEL_CSR_TIME", "PARALLEL_SIMD_CSR_TIME", "PARALLEL_HLL_TIME",
           "PARALLEL_HLL_SIMD_TIME", "SERIAL_HLL_TIME", "ROW_CSR_TIME", "WARP_CSR_TIME", "ROW_HLL_TIME",
           "WARP_HLL_TIME", "WARP_SHARED_MEMORY_CSR_TIME", "WARP_SHARED_MEMORY_HLL_TIME", "B200_CSR_TIME",
           "B200_HLL_TIME", "B200_CSR_E2E_TIME", "B200_HLL_E2E_TIME")
for _i, _name in enumerate(METRICS):
    globals()[_name] = _i
NUM_METRICS = len(METRICS)


def init_vector_at_one(v):
    N.lib().init_vector_at_one(_dp(v), len(v))
    return v


def calculate_flops(nz, time):
    return N.lib().calculate_flops(int(nz), float(time))


def computeDifferenceMetrics(ref, res, abs_tol=1e-5, rel_tol=1e-4, print_summary=False):
    ref, res = _f64(ref), _f64(res)
    d = N.lib().computeDifferenceMetrics(_dp(ref), _dp(res), len(ref), abs_tol, rel_tol, print_summary)
    return d.mean_abs_err, d.mean_rel_err, d.significant_diffs


def initialize_metrics():
    N.lib().initialize_metrics()


def cleanup_metrics():
    N.lib().cleanup_metrics()


def reset_medium_time_metrics():
    N.lib().reset_medium_time_metrics()


def update_medium_metric(metric, value):
    N.lib().update_medium_metric(int(metric), float(value))


def get_metric_value(metric):
    return N.lib().get_metric_value(int(metric))


def get_metric_min(metric):
    return N.lib().get_metric_min(int(metric))


def get_metric_median(metric):
    return N.lib().get_metric_median(int(metric))


def accumulateErrors(mean_abs_err, mean_rel_err, metric):
    d = N.DiffMetricsStruct(mean_abs_err, mean_rel_err, 0)
    N.lib().accumulateErrors(C.byref(d), int(metric))


def computeAverageErrors(metric):
    d = N.lib().computeAverageErrors(int(metric))
    return d.mean_abs_err, d.mean_rel_err


def calculate_csr_bytes(M, N_, nnz, value_bytes=8):
    return N.lib().calculate_csr_bytes(int(M), int(N_), int(nnz), value_bytes)


def calculate_hll_bytes(M, N_, slots, num_blocks, value_bytes=8):
    return N.lib().calculate_hll_bytes(int(M), int(N_), int(slots), int(num_blocks), value_bytes)


def sort_row(col_idx, values, low, high):
    """In-place reference row quicksort on int32 / float64 arrays, [low, high] inclusive."""
    assert col_idx.dtype == np.int32 and values.dtype == np.float64
    N.lib().sort_row(_ip(col_idx), _dp(values), int(low), int(high))
