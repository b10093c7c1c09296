#!/usr/bin/env python
"""Batch-size sweep of the thread-per-row kernels (csr_row_kernel, hll_row_kernel) on the stencil matrices.
The batch is read from the environment at launch time, so one process sweeps all values."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from sparsematrixvectormultiplication_b200 import device, synth  # noqa: E402
from tune import timeit  # noqa: E402

torch.cuda.set_device(0)
for name, kind, p in (("lap2d_4096", synth.SYNTH_LAP2D, 4096), ("lap3d_384", synth.SYNTH_LAP3D, 384), ("lap3d_512", synth.SYNTH_LAP3D, 512)):
    A = device.DeviceCSR.synth(kind, p)
    H = A.to_hll()
    i, hi = A.info(), H.info()
    x = torch.empty(i.N, dtype=torch.float64, device="cuda")
    device.synth_vector(x, 7)
    y = torch.empty(i.M, dtype=torch.float64, device="cuda")
    base = min(timeit(lambda: A.spmv(x, y, algo=device.ALGO_STREAM), 20, 3) for _ in range(3))
    hbase = min(timeit(lambda: H.spmv(x, y, slice_kernel=False), 20, 3) for _ in range(3))
    print(f"{name}: csr stream {base*1e3:.1f} us {i.algorithmic_bytes/base/1e6:.0f} GB/s | hll stream {hbase*1e3:.1f} us {hi.algorithmic_bytes/hbase/1e6:.0f} GB/s", flush=True)
    for b in range(1, 9):
        os.environ["SPMV_B200_ROW_BATCH"] = str(b)
        os.environ["SPMV_B200_HLL_ROW_BATCH"] = str(b)
        os.environ["SPMV_B200_VECTOR_WIDTH"] = "1"
        t = min(timeit(lambda: A.spmv(x, y, algo=device.ALGO_VECTOR), 20, 3) for _ in range(3)) 
        th = min(timeit(lambda: H.spmv(x, y, slice_kernel="rows"), 20, 3) for _ in range(3))
        print(f"{name}: batch {b}  csr_row {t*1e3:7.1f} us {i.algorithmic_bytes/t/1e6:6.0f} GB/s | hll_row {th*1e3:7.1f} us {hi.algorithmic_bytes/th/1e6:6.0f} GB/s", flush=True)
    H.close()
    A.close()
    del x, y
    torch.cuda.empty_cache()
