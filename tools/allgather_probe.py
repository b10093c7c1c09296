#!/usr/bin/env python
"""Times the in-place all-gather of the padded x of lap3d 512^3 alone (torchrun, one rank per GPU) -- run under different
NCCL_* environment settings to see what the collective itself can reach on this box."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stride = ((512 ** 3 // world + 2 * 512 * 512) + 31) // 32 * 32
xg = torch.ones(world * stride, dtype=torch.float64, device="cuda")
own = xg[rank * stride:(rank + 1) * stride]
for _ in range(5):
    dist.all_gather_into_tensor(xg, own)
torch.cuda.synchronize()
dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
reps = 20
for _ in range(reps):
    dist.all_gather_into_tensor(xg, own)
b.record()
torch.cuda.synchronize()
ms = torch.tensor([a.elapsed_time(b) / reps], device="cuda")
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    knobs = {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}
    print(f"all-gather {world} x {8 * stride / 1e6:.1f} MB: {ms.item():.3f} ms, ingress {(world - 1) * 8 * stride / ms.item() / 1e6:.0f} GB/s  {knobs}", flush=True)
dist.barrier()
dist.destroy_process_group()
