#!/usr/bin/env python
"""Time the main single-GPU workloads with whatever build SPMV_B200_LIB points at (default: the in-tree library).
Used to compare compile-time variants on ONE box:  for v in a b; do SPMV_B200_LIB=.../libspmv_$v.so python tools/ab_variants.py; done
Usage: python tools/ab_variants.py [tag] [workload ...]   workloads: lap2d uniform rmat lap3d"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from sparsematrixvectormultiplication_b200 import device, synth  # noqa: E402
from tune import timeit  # noqa: E402


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("SPMV_B200_LIB", "default")
    names = sys.argv[2:] or ["lap2d", "uniform", "rmat"]
    torch.cuda.set_device(0)
    for name in names:
        if name == "lap2d":
            A = device.DeviceCSR.synth(synth.SYNTH_LAP2D, 4096)
        elif name == "lap3d":
            A = device.DeviceCSR.synth(synth.SYNTH_LAP3D, 384)
        elif name == "uniform":
            A = device.DeviceCSR.synth(synth.SYNTH_UNIFORM, 1 << 23, 1 << 23, 32)
        elif name == "rmat":
            rp, ci, va = synth.rmat_csr_device(24, 16)
            A = device.DeviceCSR.wrap(1 << 24, 1 << 24, rp, ci, va)
        else:
            raise SystemExit(name)
        i = A.info()
        x = torch.empty(i.N, dtype=torch.float64, device="cuda")
        device.synth_vector(x, 4242)
        y = torch.empty(i.M, dtype=torch.float64, device="cuda")
        gb = i.algorithmic_bytes / 1e6
        out = []
        for label, algo in (("auto", device.ALGO_AUTO), ("stream", device.ALGO_STREAM), ("tile", device.ALGO_TILE), ("vector", device.ALGO_VECTOR), ("binned", device.ALGO_BINNED)):
            if name == "rmat" and label == "vector":
                continue
            t = min(timeit(lambda: A.spmv(x, y, algo=algo), 20, 3) for _ in range(3))
            out.append(f"{label} {t*1e3:.1f} us {gb/t:.0f} GB/s")
        if name != "rmat":
            H = A.to_hll()
            hb = H.info().algorithmic_bytes / 1e6
            for label, flag in (("hll-auto", None), ("hll-stream", False), ("hll-slice", True), ("hll-rows", "rows")):
                t = min(timeit(lambda: H.spmv(x, y, slice_kernel=flag), 20, 3) for _ in range(3))
                out.append(f"{label} {t*1e3:.1f} us {hb/t:.0f} GB/s")
            H.close()
        print(f"[{tag}] {name}: " + " | ".join(out), flush=True)
        A.close()
        del x, y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
