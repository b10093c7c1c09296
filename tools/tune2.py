#!/usr/bin/env python
"""Sweep of the knobs that trade shared memory (TMA ring) against L1 (gathers in flight) on gather-heavy matrices.
Usage: python tools/tune2.py [uniform|rmat|lap2d] ..."""
import itertools
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from sparsematrixvectormultiplication_b200 import device, synth  # noqa: E402
from tune import timeit  # noqa: E402


def build(kind):
    if kind == "lap2d":
        return device.DeviceCSR.synth(synth.SYNTH_LAP2D, 4096)
    if kind == "uniform":
        return device.DeviceCSR.synth(synth.SYNTH_UNIFORM, 1 << 23, 1 << 23, 32)
    if kind == "rmat":
        rp, ci, va = synth.rmat_csr_device(24, 16)
        A = device.DeviceCSR.wrap(1 << 24, 1 << 24, rp, ci, va)
        A._keep = (rp, ci, va)
        return A
    raise SystemExit(kind)


def main(kinds):
    torch.cuda.set_device(0)
    for kind in kinds:
        A = build(kind)
        i = A.info()
        x = torch.empty(i.N, dtype=torch.float64, device="cuda")
        device.synth_vector(x, 4242)
        y = torch.empty(i.M, dtype=torch.float64, device="cuda")
        gb = i.algorithmic_bytes / 1e6
        for D in (4608, 6144, 9216):
            A.replan(tile_items=D, long_threshold=512)
            t = timeit(lambda: A.spmv(x, y, algo=device.ALGO_TILE), 10, 2)
            print(f"{kind:8s} tile    D={D:5d}  {t*1e3:8.1f} us  {gb/t:6.0f} GB/s", flush=True)
        for per_sm, cons, stages, D in itertools.product((1, 2), (8, 12, 16, 24), (2, 3), (4608, 6144, 9216, 12288)):
            if (per_sm == 1) != (cons == 24):
                continue
            os.environ["SPMV_B200_CTAS_PER_SM"] = str(per_sm)
            os.environ["SPMV_B200_CONSUMER_WARPS"] = str(cons)
            os.environ["SPMV_B200_STAGES"] = str(stages)
            try:
                A.replan(tile_items=D, long_threshold=512)
                t = timeit(lambda: A.spmv(x, y, algo=device.ALGO_STREAM), 10, 2)
                print(f"{kind:8s} stream  ctas/sm={per_sm} consumers={cons:2d} stages={stages} D={D:5d}  {t*1e3:8.1f} us  {gb/t:6.0f} GB/s", flush=True)
            except Exception as e:
                print(f"{kind:8s} stream  ctas/sm={per_sm} consumers={cons:2d} stages={stages} D={D:5d}  failed: {e}", flush=True)
        A.close()
        del x, y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main(sys.argv[1:] or ["uniform", "rmat"])
