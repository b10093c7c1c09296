#!/usr/bin/env python
"""Does holding the head of a matrix array in the persisting L2 carve-out ACROSS products pay?  (SPMV_B200_MATRIX_PERSIST:
row_ptr of the thread-per-row CSR kernels, JA of the HLL row kernels.)  Times every workload with the policy off, then on
(the carve-out is raised on first use, so the 'off' pass runs first in a process of its own).
    python tools/persist_probe.py off|on [pct ...]"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from sparsematrixvectormultiplication_b200 import device, partition, synth  # noqa: E402


def timed(fn, reps=60, warm=8):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        best.append(a.elapsed_time(b) / reps * 1e3)
    return min(best), sorted(best)[1]


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "off"
    pcts = [int(v) for v in sys.argv[2:]] or [100]
    torch.cuda.set_device(0)
    os.environ["SPMV_B200_MATRIX_PERSIST"] = "0"
    work = []
    A2 = device.DeviceCSR.synth(synth.SYNTH_LAP2D, 4096)
    H2 = A2.to_hll()
    A2.enable_f32()
    H2.enable_f32()
    n2 = 4096 * 4096
    x = torch.ones(n2, dtype=torch.float64, device="cuda")
    y = torch.empty(n2, dtype=torch.float64, device="cuda")
    x32, y32 = x.float(), torch.empty(n2, dtype=torch.float32, device="cuda")
    work += [("lap2d csr f64", lambda: A2.spmv(x, y)), ("lap2d hll f64", lambda: H2.spmv(x, y)),
             ("lap2d csr f32", lambda: A2.spmv_f32(x32, y32)), ("lap2d hll f32", lambda: H2.spmv_f32(x32, y32))]
    lo, hi = partition.synth_partition(synth.SYNTH_LAP3D, 512, 0, 0, 8)[3]
    A3 = device.DeviceCSR.synth(synth.SYNTH_LAP3D, 512, row_begin=lo, row_end=hi)
    x3 = torch.ones(512 ** 3, dtype=torch.float64, device="cuda")
    y3 = torch.empty(hi - lo, dtype=torch.float64, device="cuda")
    p3 = torch.zeros(A3.flat_partials_count(), dtype=torch.float64, device="cuda")
    inv = torch.ones(1, dtype=torch.float64, device="cuda")
    work += [("lap3d 512 rank 3 of 8, csr flat fused", lambda: A3.spmv_fused_flat(x3.data_ptr(), y3.data_ptr(), inv_norm=inv, partials=p3)),
             ("lap3d 512 rank 3 of 8, csr plain", lambda: A3.spmv(x3, y3))]
    lo_h, hi_h = partition.hack_aligned(partition.synth_partition(synth.SYNTH_LAP3D, 512, 0, 0, 8), 512 ** 3)[3]
    A3h = device.DeviceCSR.synth(synth.SYNTH_LAP3D, 512, row_begin=lo_h, row_end=hi_h)
    H3 = A3h.to_hll()
    A3h.close()
    p3h = torch.zeros(H3.flat_partials_count(), dtype=torch.float64, device="cuda")
    y3h = torch.empty(hi_h - lo_h, dtype=torch.float64, device="cuda")
    work += [("lap3d 512 rank 3 of 8, hll flat fused", lambda: H3.spmv_fused_flat(x3.data_ptr(), y3h.data_ptr(), inv_norm=inv, partials=p3h))]
    if mode == "off":
        for name, fn in work:
            t, med = timed(fn)
            print(f"persist off      {name}: {t:8.1f} us (median of 3: {med:.1f})", flush=True)
        return
    os.environ["SPMV_B200_MATRIX_PERSIST"] = "1"
    for pct in pcts:
        os.environ["SPMV_B200_MATRIX_PERSIST_PCT"] = str(pct)
        for name, fn in work:
            t, med = timed(fn)
            print(f"persist on {pct:3d} % {name}: {t:8.1f} us (median of 3: {med:.1f})", flush=True)


if __name__ == "__main__":
    main()
