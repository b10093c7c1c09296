#!/usr/bin/env python
"""Which row kernel the plan-time timing picks for the HLL image of lap2d 4096^2 (fp64 and fp32 storage) and what it runs at."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from sparsematrixvectormultiplication_b200 import device, synth  # noqa: E402
from tune import timeit  # noqa: E402

torch.cuda.set_device(0)
A = device.DeviceCSR.synth(synth.SYNTH_LAP2D, 4096)
H = A.to_hll().enable_f32()
i, hi = A.info(), H.info()
x = torch.empty(i.N, dtype=torch.float64, device="cuda")
device.synth_vector(x, 7)
y, y2 = torch.empty(i.M, dtype=torch.float64, device="cuda"), torch.empty(i.M, dtype=torch.float64, device="cuda")
t = min(timeit(lambda: H.spmv(x, y), 50, 5) for _ in range(3))
print(f"lap2d_4096 hll f64 automatic choice: {H.row_form()}  {t*1e3:.1f} us  {hi.algorithmic_bytes/t/1e6:.0f} GB/s  {2*i.nnz/t/1e6:.0f} GFLOP/s")
A.spmv(x, y2)
print("bitwise equal to the CSR row kernel:", bool(torch.equal(y, y2)))
x32, y32 = x.float(), torch.empty(i.M, dtype=torch.float32, device="cuda")
t = min(timeit(lambda: H.spmv_f32(x32, y32), 50, 5) for _ in range(3))
print(f"lap2d_4096 hll f32 automatic choice: {H.row_form_f32()}  {t*1e3:.1f} us  {H.algorithmic_bytes_f32()/t/1e6:.0f} GB/s")
