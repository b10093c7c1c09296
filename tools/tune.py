#!/usr/bin/env python
"""Parameter sweep of the stream kernels on the GPU box (plan knobs are read from the environment at plan time).
Usage: python tools/tune.py [lap2d|lap3d|uniform|rmat] ...   -> one line per configuration on stdout."""
import itertools
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from sparsematrixvectormultiplication_b200 import device, synth  # noqa: E402


def timeit(fn, steps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def build(kind):
    if kind == "lap2d":
        return device.DeviceCSR.synth(synth.SYNTH_LAP2D, 4096)
    if kind == "lap3d":
        return device.DeviceCSR.synth(synth.SYNTH_LAP3D, 320)
    if kind == "uniform":
        return device.DeviceCSR.synth(synth.SYNTH_UNIFORM, 1 << 22, 1 << 22, 32)
    if kind == "rmat":
        rp, ci, va = synth.rmat_csr_device(22, 16)
        A = device.DeviceCSR.wrap(1 << 22, 1 << 22, rp, ci, va)
        A._rmat = (rp, ci, va)
        return A
    raise SystemExit(kind)


def main(kinds):
    torch.cuda.set_device(0)
    grid = {
        "SPMV_B200_STAGES": [2, 3],
        "SPMV_B200_CONSUMER_WARPS": [12, 16, 24],
        "D": [6144, 8192, 10240, 12288, 16384, 24576],
    }
    for kind in kinds:
        A = build(kind)
        i = A.info()
        x = torch.empty(i.N, dtype=torch.float64, device="cuda")
        device.synth_vector(x, 4242)
        y = torch.empty(i.M, dtype=torch.float64, device="cuda")
        ref = None
        for stages, cons, D in itertools.product(*grid.values()):
            os.environ["SPMV_B200_STAGES"] = str(stages)
            os.environ["SPMV_B200_CONSUMER_WARPS"] = str(cons)
            try:
                A.replan(tile_items=D, long_threshold=512)
                ms = timeit(lambda: A.spmv(x, y))
                if ref is None:
                    ref = y.clone()
                ok = bool(torch.allclose(y, ref, rtol=1e-12, atol=0))
                print(f"{kind:8s} csr  stages={stages} consumers={cons:2d} D={D:5d}  {ms*1e3:8.1f} us  {i.algorithmic_bytes/ms/1e6:7.0f} GB/s  ok={ok}", flush=True)
            except Exception as e:
                print(f"{kind:8s} csr  stages={stages} consumers={cons:2d} D={D:5d}  failed: {e}", flush=True)
        if kind in ("lap2d", "uniform"):
            for stages, cons, D in itertools.product([2, 3], [12, 16, 24], [2048, 3072, 3520, 4096, 6144]):
                os.environ["SPMV_B200_STAGES"] = str(stages)
                os.environ["SPMV_B200_CONSUMER_WARPS"] = str(cons)
                os.environ["SPMV_B200_HLL_TILE_SLOTS"] = str(D)
                try:
                    H = A.to_hll()
                    hi = H.info()
                    ms = timeit(lambda: H.spmv(x, y))
                    print(f"{kind:8s} hll  stages={stages} consumers={cons:2d} D={D:5d}  {ms*1e3:8.1f} us  {hi.algorithmic_bytes/ms/1e6:7.0f} GB/s", flush=True)
                    H.close()
                except Exception as e:
                    print(f"{kind:8s} hll  stages={stages} consumers={cons:2d} D={D:5d}  failed: {e}", flush=True)
        A.close()


if __name__ == "__main__":
    main(sys.argv[1:] or ["lap2d"])
