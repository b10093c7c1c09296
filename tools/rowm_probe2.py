#!/usr/bin/env python
"""Second probe of the fp32 row path on lap2d 4096^2 (and lap3d 256^3): what removing the FIRST of a row's three dependent
round trips buys -- hll_rowu_kernel, hack offsets by arithmetic on regular images (SPMV_B200_HLL_UNIFORM = batch, read at
launch time) -- next to the index-only forms and the plain kernels.  The run kept as
profiles/r02e_rowm_probe_third_pass.log also swept two switches that were removed afterwards because they lost: a
prefetch of the row_ptr / hack_off line of a later CTA into L2, and the offset array in the persisting L2 carve-out.
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
KEYS = ("SPMV_B200_ROW_MULTI", "SPMV_B200_HLL_UNIFORM")


def best(fn):
    return min(timeit(fn, 20, 3) for _ in range(3))


def setting(**kw):
    for k in KEYS:
        os.environ.pop(k, None)
    for k, v in kw.items():
        os.environ["SPMV_B200_" + k] = str(v)


SHAPES = (("lap2d_4096", synth.SYNTH_LAP2D, 4096), ("lap3d_256", synth.SYNTH_LAP3D, 256))
for name, kind, p in [s for s in SHAPES if not sys.argv[1:] or s[0] in sys.argv[1:]]:
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
    setting()
    H.close()
    A.close()
    del x, y, x32, y32
    torch.cuda.empty_cache()

