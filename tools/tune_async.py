#!/usr/bin/env python
"""One GPU, lap3d n^3: iteration time of the mailbox and the asynchronous fused products next to the plain product."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from sparsematrixvectormultiplication_b200 import device, synth  # noqa: E402
from sparsematrixvectormultiplication_b200.distributed import AsyncPowerIteration, FusedPowerIteration  # noqa: E402
from tune import timeit  # noqa: E402

torch.cuda.set_device(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
F = FusedPowerIteration(synth.SYNTH_LAP3D, n, mailbox=True)
i = F.A.info()
gb = i.algorithmic_bytes / 1e6
x = torch.ones(i.N, dtype=torch.float64, device="cuda")
y = torch.empty(i.M, dtype=torch.float64, device="cuda")
a = min(timeit(lambda: F.A.spmv(x, y), 20, 3) for _ in range(3))
b = min(timeit(F.step, 20, 3) for _ in range(3))
print(f"lap3d_{n}: plain {a*1e3:.1f} us {gb/a:.0f} GB/s (auto {device.ALGO_NAMES[i.auto_algo]}, batch {i.row_batch}) | mailbox step {b*1e3:.1f} us {gb/b:.0f} GB/s", flush=True)
F.close()
del F
G = AsyncPowerIteration(synth.SYNTH_LAP3D, n)
c = min(timeit(G.step, 20, 3) for _ in range(3))
print(f"lap3d_{n}: async step {c*1e3:.1f} us {gb/c:.0f} GB/s", flush=True)
G.close()
