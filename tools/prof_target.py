#!/usr/bin/env python
"""Small, deterministic launch sequence for ncu: for each named workload build the matrix on the device
and run the product a few times.  Usage: python tools/prof_target.py lap2d_csr lap2d_hll uniform_csr ...
Every workload launches exactly REPS products per kernel so that -s/-c can pick one warm launch."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from sparsematrixvectormultiplication_b200 import device, synth  # noqa: E402

REPS = 3


def vec(n, seed=4242):
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    device.synth_vector(x, seed)
    return x


def main(names):
    torch.cuda.set_device(0)
    for name in names:
        kind, fmt = name.rsplit("_", 1)
        if kind == "lap2d":
            A = device.DeviceCSR.synth(synth.SYNTH_LAP2D, 4096)
        elif kind == "lap3d":
            A = device.DeviceCSR.synth(synth.SYNTH_LAP3D, 256)
        elif kind == "uniform":
            A = device.DeviceCSR.synth(synth.SYNTH_UNIFORM, 1 << 23, 1 << 23, 32)
        elif kind == "rmat":
            rp, ci, va = synth.rmat_csr_device(24, 16)
            A = device.DeviceCSR.wrap(1 << 24, 1 << 24, rp, ci, va)
        else:
            raise SystemExit(f"unknown workload {name}")
        info = A.info()
        x, y = vec(info.N), torch.empty(info.M, dtype=torch.float64, device="cuda")
        if fmt == "csr":
            fn = lambda: A.spmv(x, y)
        elif fmt == "vec":
            fn = lambda: A.spmv(x, y, algo=device.ALGO_VECTOR)
        elif fmt == "hll":
            H = A.to_hll()
            fn = lambda: H.spmv(x, y)
        elif fmt == "csr32":      # fp32 storage, fp64 arithmetic
            A.enable_f32()
            x32, y32 = x.float(), torch.empty(info.M, dtype=torch.float32, device="cuda")
            fn = lambda: A.spmv_f32(x32, y32)
        elif fmt == "hll32":
            H = A.to_hll()
            H.enable_f32()
            x32, y32 = x.float(), torch.empty(info.M, dtype=torch.float32, device="cuda")
            fn = lambda: H.spmv_f32(x32, y32)
        else:
            raise SystemExit(f"unknown format in {name}")
        fn()                                  # warm (plan-time tuning of the fp32 paths happened in enable_f32)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("prof")    # ncu --nvtx --nvtx-include "prof/" profiles only these launches
        for _ in range(REPS):
            fn()
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        print(name, "done", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["lap2d_csr"])
