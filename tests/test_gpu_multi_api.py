"""The single-process multi-GPU C entry points (spmv_b200_multi_*) on GENERAL matrices handed over as host CSR arrays
(the reference's CSRMatrix fields), bound with ctypes: the row-partitioned product against the serial oracle, and the power
iteration in all three exchange modes (MAILBOX, NCCL ALLGATHER, ALLGATHER_PEER) against the oracle's power iteration, for
  * short rows (<= 12 nonzeros: the two-launch flat form),
  * medium rows (<= 40: one fused launch of the stream kernel),
  * a matrix with a 3000-nonzero row (MAILBOX is refused with a clear error, ALLGATHER works),
CSR and HLL, on every GPU count the box offers (1 on the driver's test box; 2+ under gpurun --gpus N)."""
import ctypes as C

import numpy as np
import pytest

from sparsematrixvectormultiplication_b200 import _native as N

pytestmark = pytest.mark.gpu
TOL = 1e-12
CSR, HLL = 0, 1
MAILBOX, ALLGATHER, ALLGATHER_PEER = 0, 1, 2


def random_square(rng, M, max_len, long_row=0):
    lengths = rng.integers(1, max_len + 1, size=M)
    if long_row:
        lengths[M // 3] = long_row
    rp = np.zeros(M + 1, np.int32)
    np.cumsum(lengths, out=rp[1:])
    ci = np.empty(rp[-1], np.int32)
    for r in range(M):
        ci[rp[r]:rp[r + 1]] = np.sort(rng.choice(M, size=lengths[r], replace=False))
    va = rng.uniform(0.1, 1.0, size=rp[-1])
    return rp, ci, va


def gpu_counts():
    import torch
    n = torch.cuda.device_count()
    return sorted({1, min(2, n), min(4, n), n})


class Multi:
    def __init__(self, ngpus, fmt, rp, ci, va):
        self.h = C.c_void_p()
        M = len(rp) - 1
        N.check(N.lib().spmv_b200_multi_init_csr(ngpus, fmt, M, M, int(rp[-1]), rp.ctypes.data_as(C.c_void_p),
                                                 ci.ctypes.data_as(C.c_void_p), va.ctypes.data_as(C.c_void_p), C.byref(self.h)))
        self.M = M

    def info(self):
        i = N.MultiInfo()
        N.check(N.lib().spmv_b200_multi_info(self.h, C.byref(i)))
        return i

    def spmv(self, x):
        y = np.full(self.M, np.nan)
        N.check(N.lib().spmv_b200_multi_spmv(self.h, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p)))
        return y

    def iterate(self, iters, mode):
        lam, ms = C.c_double(), C.c_double()
        rc = N.lib().spmv_b200_multi_iterate(self.h, iters, mode, C.byref(lam), C.byref(ms))
        return rc, lam.value

    def reset(self):
        N.check(N.lib().spmv_b200_multi_reset(self.h, None))

    def x(self):
        out = np.full(self.M, np.nan)
        N.check(N.lib().spmv_b200_multi_get_x(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def close(self):
        N.lib().spmv_b200_multi_free(self.h)


@pytest.mark.parametrize("fmt", [CSR, HLL])
@pytest.mark.parametrize("shape", ["short", "medium", "long"])
def test_multi_entry_points_on_general_matrices(checker, port, fmt, shape):
    rng = np.random.default_rng({"short": 1, "medium": 2, "long": 3}[shape])
    M = 6000
    rp, ci, va = random_square(rng, M, {"short": 12, "medium": 40, "long": 20}[shape], long_row=3000 if shape == "long" else 0)
    x = rng.uniform(0.5, 1.5, size=M)
    y_ref = checker.spmv_csr_serial(rp, ci, va, x)
    iters = 8
    x_ref, _, lam_ref = port.power_iteration(rp, ci, va, np.ones(M), iters)
    for ngpus in gpu_counts():
        A = Multi(ngpus, fmt, rp, ci, va)
        try:
            info = A.info()
            assert info.M == M and info.nnz == rp[-1] and 1 <= info.ngpus <= ngpus
            assert info.row_begin[0] == 0 and info.row_end[info.ngpus - 1] == M
            assert sum(info.nnz_part[g] for g in range(info.ngpus)) == rp[-1]
            y = A.spmv(x)
            assert np.all(np.abs(y - y_ref) <= TOL * y_ref), (shape, fmt, ngpus)     # all terms positive
            for mode in (MAILBOX, ALLGATHER, ALLGATHER_PEER):
                A.reset()
                rc, lam = A.iterate(iters, mode)
                if shape == "long" and mode == MAILBOX:
                    assert info.fused_ok == 0 and rc == -1
                    assert "long-row" in N.last_error()
                    continue
                assert rc == 0, N.last_error()
                assert abs(lam - lam_ref[-1]) <= 1e-11 * lam_ref[-1], (shape, fmt, ngpus, mode, lam, lam_ref[-1])
                assert np.max(np.abs(A.x() - x_ref)) <= 1e-11 * np.max(np.abs(x_ref)), (shape, fmt, ngpus, mode)
        finally:
            A.close()
