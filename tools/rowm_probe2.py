#!/usr/bin/env python
"""Second probe of the fp32 row path on lap2d 4096^2 (and lap3d 256^3): what removing or shortening the FIRST of a row's
three dependent round trips buys.  (a) hll_rowu_kernel: hack offsets by arithmetic (regular images), (b) the offset /
row_ptr line of a later CTA prefetched into L2 (SPMV_B200_ROW_PREFETCH = CTAs ahead), (c) the offset array held in the
persisting L2 carve-out.  Everything is an environment variable read at launch time, so one process sweeps all of it.
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from sparsematrixvectormultiplication_b200 import device, synth  # noqa: E402
from tune import timeit  # noqa: E402

torch.cuda.set_device(0)
os.environ["SPMV_B200_AUTOTUNE"] = "0"
KEYS = ("SPMV_B200_ROW_MULTI", "SPMV_B200_HLL_UNIFORM", "SPMV_B200_ROW_PREFETCH", "SPMV_B200_MATRIX_PERSIST", "SPMV_B200_HACKOFF_PERSIST")


def best(fn):
    return min(timeit(fn, 20, 3) for _ in range(3))


def setting(**kw):
    for k in KEYS:
        os.environ.pop(k, None)
    for k, v in kw.items():
        os.environ["SPMV_B200_" + k] = str(v)


for name, kind, p in (("lap2d_4096", synth.SYNTH_LAP2D, 4096), ("lap3d_256", synth.SYNTH_LAP3D, 256)):
    A = device.DeviceCSR.synth(kind, p).enable_f32()
    H = A.to_hll().enable_f32()
    i = A.info()
    x = torch.empty(i.N, dtype=torch.float64, device="cuda")
    device.synth_vector(x, 7)
    y = torch.empty(i.M, dtype=torch.float64, device="cuda")
    x32, y32 = x.float(), torch.empty(i.M, dtype=torch.float32, device="cuda")
    csr32 = lambda: A.spmv_f32(x32, y32, algo=device.ALGO_ROW)
    hll32 = lambda: H.spmv_f32(x32, y32)
    hll64 = lambda: H.spmv(x, y, slice_kernel="rows")
    forms = {"csr": device.row_forms(device.FORMAT_CSR), "hll": device.row_forms(device.FORMAT_HLL)}
    # index of the one-row index-only forms with batch 5 / 7
    pick = {fmt: {b: 1 + forms[fmt].index((1, b, 8)) for b in (4, 5, 7) if (1, b, 8) in forms[fmt]} for fmt in forms}
    b0 = 5 if kind == synth.SYNTH_LAP2D else 7
    setting()
    print(f"{name} plain (default batch): csr f32 {best(csr32)*1e3:6.1f} us | hll f32 {best(hll32)*1e3:6.1f} us | hll f64 {best(hll64)*1e3:6.1f} us", flush=True)
    for b in sorted(pick["hll"]):
        setting(ROW_MULTI=pick["hll"][b])
        t_h = best(hll32)
        setting(ROW_MULTI=pick["csr"][b])
        t_c = best(csr32)
        print(f"{name} index-only form batch {b}: csr f32 {t_c*1e3:6.1f} us | hll f32 {t_h*1e3:6.1f} us", flush=True)
    for b in (4, 5, 6, 7):
        setting(HLL_UNIFORM=b)
        print(f"{name} hll_rowu batch {b}: hll f32 {best(hll32)*1e3:6.1f} us {H.algorithmic_bytes_f32()/best(hll32)/1e6:5.0f} GB/s | hll f64 {best(hll64)*1e3:6.1f} us", flush=True)
    for ahead in (148, 592, 1184, 2368, 4736, 9472):
        setting(ROW_MULTI=pick["hll"][b0], ROW_PREFETCH=ahead)
        t_h = best(hll32)
        setting(ROW_MULTI=pick["csr"][b0], ROW_PREFETCH=ahead)
        t_c = best(csr32)
        print(f"{name} prefetch {ahead} CTAs ahead (index-only form batch {b0}): csr f32 {t_c*1e3:6.1f} us | hll f32 {t_h*1e3:6.1f} us", flush=True)
    setting()
    H.close()
    A.close()
    del x, y, x32, y32
    torch.cuda.empty_cache()

# last, because raising the persisting carve-out is a per-device setting that stays for the rest of the process
A = device.DeviceCSR.synth(synth.SYNTH_LAP2D, 4096).enable_f32()
H = A.to_hll().enable_f32()
i = A.info()
x32 = torch.ones(i.N, dtype=torch.float32, device="cuda")
y32 = torch.empty(i.M, dtype=torch.float32, device="cuda")
form = 1 + device.row_forms(device.FORMAT_HLL).index((1, 5, 8))
setting(ROW_MULTI=form)
t0 = best(lambda: H.spmv_f32(x32, y32))
setting(ROW_MULTI=form, MATRIX_PERSIST=1, HACKOFF_PERSIST=1)
t1 = best(lambda: H.spmv_f32(x32, y32))
print(f"lap2d_4096 hack_off in the persisting carve-out (index-only form batch 5): hll f32 {t1*1e3:6.1f} us against {t0*1e3:6.1f} us without", flush=True)
setting()
