#!/usr/bin/env python
"""Every multi-row form of the row kernels (csr_rowm_kernel / hll_rowm_kernel, forced through SPMV_B200_ROW_MULTI, which
is read at launch time) against the one-row kernels at each batch, on the stencil matrices, fp32 and fp64 storage.
Prints one line per form: time of one product (minimum of three 20-launch means) and GB/s of algorithmic bytes.

    python tools/rowm_probe.py [lap2d_4096 lap3d_256 ...]
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from sparsematrixvectormultiplication_b200 import device, synth  # noqa: E402
from tune import timeit  # noqa: E402

SHAPES = {"lap2d_4096": (synth.SYNTH_LAP2D, 4096), "lap3d_256": (synth.SYNTH_LAP3D, 256), "lap3d_384": (synth.SYNTH_LAP3D, 384)}
torch.cuda.set_device(0)
os.environ["SPMV_B200_AUTOTUNE"] = "0"   # forms are forced below; no plan-time timing


def best(fn):
    return min(timeit(fn, 20, 3) for _ in range(3))


for name in (sys.argv[1:] or ["lap2d_4096", "lap3d_256"]):
    kind, p = SHAPES[name]
    A = device.DeviceCSR.synth(kind, p).enable_f32()
    H = A.to_hll().enable_f32()
    i = A.info()
    x = torch.empty(i.N, dtype=torch.float64, device="cuda")
    device.synth_vector(x, 7)
    y = torch.empty(i.M, dtype=torch.float64, device="cuda")
    x32, y32 = x.float(), torch.empty(i.M, dtype=torch.float32, device="cuda")
    legs = (("csr f32", lambda: A.spmv_f32(x32, y32, algo=device.ALGO_ROW), A.algorithmic_bytes_f32()),
            ("hll f32", lambda: H.spmv_f32(x32, y32), H.algorithmic_bytes_f32()),
            ("csr f64", lambda: A.spmv(x, y, algo=device.ALGO_ROW), i.algorithmic_bytes),
            ("hll f64", lambda: H.spmv(x, y, slice_kernel="rows"), H.info().algorithmic_bytes))
    os.environ.pop("SPMV_B200_ROW_MULTI", None)
    for b in range(2, 8):
        os.environ["SPMV_B200_ROW_BATCH"] = os.environ["SPMV_B200_HLL_ROW_BATCH"] = str(b)
        # the batch variables are read at plan time: a fresh pair of handles per batch
        A1 = device.DeviceCSR.synth(kind, p).enable_f32()
        H1 = A1.to_hll().enable_f32()
        one = (("csr f32", lambda: A1.spmv_f32(x32, y32, algo=device.ALGO_ROW), A.algorithmic_bytes_f32()),
               ("hll f32", lambda: H1.spmv_f32(x32, y32), H.algorithmic_bytes_f32()),
               ("csr f64", lambda: A1.spmv(x, y, algo=device.ALGO_ROW), i.algorithmic_bytes),
               ("hll f64", lambda: H1.spmv(x, y, slice_kernel="rows"), H.info().algorithmic_bytes))
        times = [(leg, best(fn), nbytes) for leg, fn, nbytes in one]
        print(f"{name} one-row batch {b}: " + " | ".join(f"{leg} {t*1e3:6.1f} us {nbytes/t/1e6:5.0f} GB/s" for leg, t, nbytes in times), flush=True)
        H1.close()
        A1.close()
    os.environ.pop("SPMV_B200_ROW_BATCH", None)
    os.environ.pop("SPMV_B200_HLL_ROW_BATCH", None)
    forms = {"csr": device.row_forms(device.FORMAT_CSR), "hll": device.row_forms(device.FORMAT_HLL)}
    for k in range(max(len(forms["csr"]), len(forms["hll"]))):
        os.environ["SPMV_B200_ROW_MULTI"] = str(k + 1)
        cells = []
        for leg, fn, nbytes in legs:
            f = forms[leg[:3]]
            if k >= len(f):
                continue
            t = best(fn)
            cells.append(f"{leg} {f[k]} {t*1e3:6.1f} us {nbytes/t/1e6:5.0f} GB/s")
        print(f"{name} form {k}: " + " | ".join(cells), flush=True)
    os.environ.pop("SPMV_B200_ROW_MULTI", None)
    H.close()
    A.close()
    del x, y, x32, y32
    torch.cuda.empty_cache()
