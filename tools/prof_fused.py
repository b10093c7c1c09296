#!/usr/bin/env python
"""ncu target: the iterated product on ONE GPU with exactly the rows of rank 0 of the 8-GPU run (the first part of the
nnz-balanced partition of lap3d 512^3): a few iterations of the two-launch form (default) or the one-launch mailbox form.
    python tools/prof_fused.py [split|mailbox] [csr|hll] [parts=8]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from sparsematrixvectormultiplication_b200 import partition, synth  # noqa: E402
from sparsematrixvectormultiplication_b200.distributed import FusedPowerIteration  # noqa: E402

form = sys.argv[1] if len(sys.argv) > 1 else "split"
fmt = sys.argv[2] if len(sys.argv) > 2 else "csr"
parts = int(sys.argv[3]) if len(sys.argv) > 3 else 8
torch.cuda.set_device(0)
first = partition.synth_partition(synth.SYNTH_LAP3D, 512, 0, 0, parts)[0]
if fmt == "hll":
    first = (first[0], first[1] // 32 * 32)
F = FusedPowerIteration(synth.SYNTH_LAP3D, 512, parts=[first], single=True, fmt=fmt, split=form == "split", mailbox=form != "split")
for _ in range(6):
    F.step()
torch.cuda.synchronize()
i = F.A.info()
print("rows", F.rows, "lambda", F.eigenvalue_estimate(), "algorithmic bytes of the local product", F.algorithmic_bytes_local,
      "plan", {"fused_batch": i.fused_batch, "flat_batch": i.flat_batch}, flush=True)
F.close()
