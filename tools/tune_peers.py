#!/usr/bin/env python
"""One GPU: cost of the peer-store part of the fused product.  'local' = the peer range points at a local scratch buffer
(stores are cheap: what remains is the per-row range check); 'empty' = an empty range (check only, no store)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from sparsematrixvectormultiplication_b200 import _native as N  # noqa: E402
from sparsematrixvectormultiplication_b200 import device, synth  # noqa: E402
from tune import timeit  # noqa: E402

torch.cuda.set_device(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
A = device.DeviceCSR.synth(synth.SYNTH_LAP3D, n)
i = A.info()
x = torch.ones(i.N, dtype=torch.float64, device="cuda")
y = torch.empty(i.M, dtype=torch.float64, device="cuda")
scratch = torch.empty(i.M, dtype=torch.float64, device="cuda")
part = torch.zeros(A.partials_count(), dtype=torch.float64, device="cuda")
gb = i.algorithmic_bytes / 1e6


def peers(lo, hi, count=1):
    ps = N.Peers()
    ps.count = count
    for k in range(count):
        ps.dst[k] = scratch.data_ptr()
        ps.lo[k], ps.hi[k] = lo, hi
    return ps


cases = {"no peers": None, "empty range": peers(0, 0), "last plane -> local buffer": peers(i.M - n * n, i.M),
         "two ranges (first + last plane)": peers(0, n * n, 1)}
two = peers(0, n * n, 2)
two.lo[1], two.hi[1] = i.M - n * n, i.M
cases["two ranges (first + last plane)"] = two
order = list(cases.items())
for name, ps in order + order[::-1]:
    t = min(timeit(lambda: A.spmv_fused(x, y, partials=part, peers=ps), 20, 3) for _ in range(3))
    print(f"lap3d_{n} fused, {name:34s}: {t*1e3:7.1f} us {gb/t:6.0f} GB/s", flush=True)
