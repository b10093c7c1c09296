#!/usr/bin/env python
"""A/B timing on one box: plain vs fused CSR stream kernel, repeated, lap2d 4096^2 and lap3d 384^3."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from sparsematrixvectormultiplication_b200 import device, synth  # noqa: E402
from tune import timeit  # noqa: E402

torch.cuda.set_device(0)
for name, A in (("lap2d_4096", device.DeviceCSR.synth(synth.SYNTH_LAP2D, 4096)), ("lap3d_384", device.DeviceCSR.synth(synth.SYNTH_LAP3D, 384))):
    i = A.info()
    x = torch.empty(i.N, dtype=torch.float64, device="cuda")
    device.synth_vector(x, 1)
    y = torch.empty(i.M, dtype=torch.float64, device="cuda")
    part = torch.zeros(A.partials_count(), dtype=torch.float64, device="cuda")
    ss = torch.ones(1, dtype=torch.float64, device="cuda")
    for rep in range(4):
        a = timeit(lambda: A.spmv(x, y), 50, 5)
        b = timeit(lambda: A.spmv_fused(x, y), 50, 5)
        c = timeit(lambda: A.spmv_fused(x, y, prev_sumsq=ss, partials=part), 50, 5)
        gb = i.algorithmic_bytes / 1e6
        print(f"{name} rep {rep}: plain {a*1e3:.1f} us {gb/a:.0f} GB/s | fused(no-op) {b*1e3:.1f} us {gb/b:.0f} GB/s | fused(scale+sumsq) {c*1e3:.1f} us {gb/c:.0f} GB/s", flush=True)
