#!/usr/bin/env python
"""ncu target: the fused iterated product on ONE GPU with the row count of one rank of the 8-GPU run
(lap3d 256^3 = 16.8 M rows = 512^3 / 8): a few mailbox iterations."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from sparsematrixvectormultiplication_b200 import synth  # noqa: E402
from sparsematrixvectormultiplication_b200.distributed import FusedPowerIteration  # noqa: E402

torch.cuda.set_device(0)
F = FusedPowerIteration(synth.SYNTH_LAP3D, int(sys.argv[1]) if len(sys.argv) > 1 else 256, mailbox=True)
for _ in range(6):
    F.step()
torch.cuda.synchronize()
print("lambda", F.eigenvalue_estimate(), "bytes", F.algorithmic_bytes_local, flush=True)
F.close()
