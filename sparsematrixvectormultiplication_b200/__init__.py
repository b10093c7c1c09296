"""sparsematrixvectormultiplication_b200 -- B200-native (sm_100a) CSR / HLL SpMV engine behind the
host API of MarcoLor01/SparseMatrixVectorMultiplication.

  host      drop-in host API (parser, CSR / HLL builders, partitioners, harness); products on the GPU
  device    resident matrices + kernels through the C-ABI of include/spmv_b200.h
  synth     synthetic matrices of the BASELINE.json configs (device generators + numpy twins)
  partition nnz-balanced row partition over GPUs (the reference's greedy rule)
  distributed  row-partitioned iterated product (power method) over torch.distributed / NCCL

All compute lives in libspmv_b200.so (hand-written CUDA for sm_100a).  There is no CPU fallback.
"""
from ._native import LIB_PATH, NativeLibraryMissing, SpmvError, build, lib  # noqa: F401

__all__ = ["LIB_PATH", "NativeLibraryMissing", "SpmvError", "build", "lib"]
