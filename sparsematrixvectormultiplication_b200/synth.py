"""Synthetic matrices of the BASELINE.json configs.

Every generator exists twice with bit-identical output:
  * on the device, inside libspmv_b200.so (csrc/cuda/synth.cu)  -- what bench.py uses at full size;
  * here in numpy -- what the CPU oracle consumes in the parity tests and the CPU baseline.
tests/test_synth.py checks the closed-form row offsets against these numpy twins on the CPU and
tests/test_gpu_parity.py checks device output == numpy output on the GPU.

The reference itself has no generator for these shapes (src/matrix_generator.py only writes 10x10
files), so the definitions are this repository's: sorted, duplicate-free rows by construction, so
the CSR that the reference's convert_in_csr would build from the same COO is exactly these arrays.
"""
from __future__ import annotations

import numpy as np

from . import _native as N

SYNTH_LAP2D, SYNTH_LAP3D, SYNTH_UNIFORM = 1, 2, 3
DEFAULT_SEED = 0x5EED

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_K1 = np.uint64(0x9E3779B97F4A7C15)
_K2 = np.uint64(0xD1B54A32D192ED03)
_VALUE_SALT = 0x5851F42D4C957F2D


def mix64(z):
    z = np.asarray(z, np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def hash3(seed, a, b):
    with np.errstate(over="ignore"):
        return mix64(np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + np.asarray(a, np.uint64) * _K1 + np.asarray(b, np.uint64) * _K2)


def unit_interval(h):
    """(0, 1] from the top 53 bits."""
    return ((np.asarray(h, np.uint64) >> np.uint64(11)) + np.uint64(1)).astype(np.float64) * (1.0 / 9007199254740992.0)


def hash_vector(n, seed=DEFAULT_SEED, begin=0):
    """x_i in (0,1]; twin of spmv_b200_synth_vector."""
    return unit_interval(hash3(seed, np.arange(begin, begin + n, dtype=np.uint64), np.uint64(0x7E)))


def row_offset(kind, p0, p1=0, p2=0, row=0) -> int:
    """nnz in rows [0,row) -- closed form from the C library (no device needed)."""
    return int(N.lib().spmv_b200_synth_row_offset(kind, int(p0), int(p1), int(p2), int(row)))


def _stencil_csr(rows, cand_cols, cand_vals, valid):
    counts = valid.sum(axis=1)
    row_ptr = np.zeros(len(rows) + 1, np.int64)
    np.cumsum(counts, out=row_ptr[1:])
    return row_ptr.astype(np.int32), cand_cols[valid].astype(np.int32), cand_vals[valid].astype(np.float64)


def lap2d_csr(n, row_begin=0, row_end=None):
    """5-point Laplacian on an n x n grid, row r = i*n + j, values -1,-1,4,-1,-1."""
    if row_end is None:
        row_end = n * n
    r = np.arange(row_begin, row_end, dtype=np.int64)
    i, j = r // n, r % n
    cols = np.stack([r - n, r - 1, r, r + 1, r + n], axis=1)
    vals = np.broadcast_to(np.array([-1.0, -1.0, 4.0, -1.0, -1.0]), cols.shape)
    valid = np.stack([i > 0, j > 0, np.ones_like(i, bool), j < n - 1, i < n - 1], axis=1)
    return _stencil_csr(r, cols, vals, valid)


def lap3d_csr(n, row_begin=0, row_end=None):
    """7-point Laplacian on an n^3 grid, row r = (i*n + j)*n + k, values -1 x6, 6 on the diagonal."""
    if row_end is None:
        row_end = n ** 3
    r = np.arange(row_begin, row_end, dtype=np.int64)
    n2 = n * n
    i, j, k = r // n2, (r // n) % n, r % n
    cols = np.stack([r - n2, r - n, r - 1, r, r + 1, r + n, r + n2], axis=1)
    vals = np.broadcast_to(np.array([-1.0, -1.0, -1.0, 6.0, -1.0, -1.0, -1.0]), cols.shape)
    valid = np.stack([i > 0, j > 0, k > 0, np.ones_like(i, bool), k < n - 1, j < n - 1, i < n - 1], axis=1)
    return _stencil_csr(r, cols, vals, valid)


def uniform_csr(M, N_, k=32, seed=DEFAULT_SEED, row_begin=0, row_end=None):
    """Exactly k nonzeros per row, entry e in column stratum e: e*(N/k) + hash % (N/k)."""
    if row_end is None:
        row_end = M
    rows = row_end - row_begin
    r = np.repeat(np.arange(row_begin, row_end, dtype=np.uint64), k)
    e = np.tile(np.arange(k, dtype=np.uint64), rows)
    stride = np.uint64(N_ // k)
    col = (e * stride + hash3(seed, r, e) % stride).astype(np.int32)
    val = unit_interval(hash3(seed + _VALUE_SALT, r, e))
    row_ptr = (np.arange(rows + 1, dtype=np.int64) * k).astype(np.int32)
    return row_ptr, col, val


def rmat_edges(scale, edges, seed=DEFAULT_SEED, a=0.57, b=0.19, c=0.19):
    """R-MAT edge list (row, col) of a 2^scale square matrix; one hash per level per edge."""
    eid = np.arange(edges, dtype=np.uint64)
    row = np.zeros(edges, np.int64)
    col = np.zeros(edges, np.int64)
    for level in range(scale):
        u = unit_interval(hash3(seed, eid, np.uint64(level)))
        right = ((u > a) & (u <= a + b)) | (u > a + b + c)   # quadrants b and d
        down = u > a + b                                      # quadrants c and d
        row = (row << 1) | down
        col = (col << 1) | right
    return row, col


def rmat_csr(scale, edge_factor=16, seed=DEFAULT_SEED, plant_dense_row=0):
    """R-MAT (0.57,0.19,0.19,0.05), duplicates removed, rows sorted, values hash(row,col) in (0,1].
    plant_dense_row > 0 adds that many distinct columns to row 0 (guarantees a heavy row at small scale)."""
    n = 1 << scale
    row, col = rmat_edges(scale, edge_factor * n, seed)
    if plant_dense_row:
        extra = np.arange(plant_dense_row, dtype=np.int64) * max(n // plant_dense_row, 1) % n
        row = np.concatenate([row, np.zeros(len(extra), np.int64)])
        col = np.concatenate([col, extra])
    key = np.unique((row << 32) | col)
    row, col = key >> 32, key & 0xFFFFFFFF
    row_ptr = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(row, minlength=n), out=row_ptr[1:])
    val = unit_interval(hash3(seed + _VALUE_SALT, row.astype(np.uint64), col.astype(np.uint64)))
    return row_ptr.astype(np.int32), col.astype(np.int32), val


def rmat_csr_device(scale, edge_factor=16, seed=DEFAULT_SEED, plant_dense_row=0, device="cuda"):
    """Same matrix built with torch ops on the GPU (generation is setup plumbing, not the hot path).
    Returns int32/int32/float64 CUDA tensors (row_ptr, col_idx, values)."""
    import torch
    n = 1 << scale
    edges = edge_factor * n
    def s64(v):  # python int -> the same 64 bits as a signed value (torch has no uint64 arithmetic)
        v &= 0xFFFFFFFFFFFFFFFF
        return v - (1 << 64) if v >= 1 << 63 else v

    M1, M2, K1, K2 = (s64(int(v)) for v in (_M1, _M2, _K1, _K2))

    def _lsr(z, s):  # logical shift right on int64
        return (z >> s) & ((1 << (64 - s)) - 1)

    def mix(z):
        z = (z ^ _lsr(z, 30)) * M1
        z = (z ^ _lsr(z, 27)) * M2
        return z ^ _lsr(z, 31)

    def h3(sd, a_, b_):
        b_term = b_ * K2 if isinstance(b_, torch.Tensor) else s64(int(b_) * K2)
        return mix(a_ * K1 + b_term + s64(sd))

    def unit(h):
        return (_lsr(h, 11) + 1).to(torch.float64) * (1.0 / 9007199254740992.0)

    a, b, c = 0.57, 0.19, 0.19
    row = torch.zeros(edges, dtype=torch.int64, device=device)
    col = torch.zeros(edges, dtype=torch.int64, device=device)
    chunk = 1 << 26
    for s in range(0, edges, chunk):
        eid = torch.arange(s, min(s + chunk, edges), dtype=torch.int64, device=device)
        rr = torch.zeros_like(eid)
        cc = torch.zeros_like(eid)
        for level in range(scale):
            u = unit(h3(seed, eid, level))
            right = ((u > a) & (u <= a + b)) | (u > a + b + c)
            down = u > a + b
            rr = (rr << 1) | down.to(torch.int64)
            cc = (cc << 1) | right.to(torch.int64)
        row[s:s + len(eid)] = rr
        col[s:s + len(eid)] = cc
    if plant_dense_row:
        extra = torch.arange(plant_dense_row, dtype=torch.int64, device=device) * max(n // plant_dense_row, 1) % n
        row = torch.cat([row, torch.zeros_like(extra)])
        col = torch.cat([col, extra])
    key = torch.unique((row << 32) | col)  # sorted
    del row, col
    r, cidx = key >> 32, key & 0xFFFFFFFF
    del key
    counts = torch.bincount(r, minlength=n)
    row_ptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    torch.cumsum(counts, 0, out=row_ptr[1:])
    val = unit(h3(seed + _VALUE_SALT, r, cidx))
    return row_ptr.to(torch.int32), cidx.to(torch.int32), val
