#!/usr/bin/env python
"""Plain vs fused product per row-kernel batch (SPMV_B200_ROW_BATCH is read at plan time)."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from sparsematrixvectormultiplication_b200 import device, synth  # noqa: E402
from tune import timeit  # noqa: E402

torch.cuda.set_device(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
for batch in (2, 5):
    os.environ["SPMV_B200_ROW_BATCH"] = str(batch)
    os.environ["SPMV_B200_FUSED_BATCH"] = str(batch)
    A = device.DeviceCSR.synth(synth.SYNTH_LAP3D, n)
    i = A.info()
    x = torch.empty(i.N, dtype=torch.float64, device="cuda")
    device.synth_vector(x, 1)
    y = torch.empty(i.M, dtype=torch.float64, device="cuda")
    part = torch.zeros(A.partials_count(), dtype=torch.float64, device="cuda")
    ss = torch.ones(1, dtype=torch.float64, device="cuda")
    gb = i.algorithmic_bytes / 1e6
    a = min(timeit(lambda: A.spmv(x, y), 20, 3) for _ in range(3))
    b = min(timeit(lambda: A.spmv_fused(x, y, prev_sumsq=ss, partials=part), 20, 3) for _ in range(3))
    c = min(timeit(lambda: A.spmv(x, y, algo=device.ALGO_STREAM), 20, 3) for _ in range(3))
    d = min(timeit(lambda: A.spmv_fused(x, y), 20, 3) for _ in range(3))
    e = min(timeit(lambda: A.spmv_fused(x, y, partials=part), 20, 3) for _ in range(3))
    print(f"   fused without scale and partials {d*1e3:.1f} us {gb/d:.0f} GB/s | fused with partials only {e*1e3:.1f} us {gb/e:.0f} GB/s")
    print(f"lap3d_{n} batch {batch or 'auto'} -> auto_algo {device.ALGO_NAMES[i.auto_algo]} row_batch {i.row_batch}: plain {a*1e3:.1f} us {gb/a:.0f} GB/s | fused {b*1e3:.1f} us {gb/b:.0f} GB/s | stream plain {c*1e3:.1f} us {gb/c:.0f} GB/s", flush=True)
    A.close()
