#!/usr/bin/env python
"""The slice push of the peer all-gather alone, under torchrun (one rank per GPU): spmv_b200_vec_push for several grid
sizes and unroll factors against one cudaMemcpyAsync per peer.  Every rank pushes `mb` MiB to every other rank at the
same time; GB/s = bytes a rank SENDS (= receives) / time, max over ranks.
    python -m torch.distributed.run --nproc-per-node N tools/push_probe.py [mib per slice]"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from sparsematrixvectormultiplication_b200 import device  # noqa: E402
from sparsematrixvectormultiplication_b200.distributed import _CudaView  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cu = torch.device("cuda", local)
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    n = mib * (1 << 20) // 8
    buf = device.PeerBuffer(8 * n * world)
    handles = [None] * world
    dist.all_gather_object(handles, buf.handle_bytes())
    src = buf.as_tensor()[rank * n: (rank + 1) * n]
    src.fill_(float(rank + 1))
    peers = [(rank + s) % world for s in range(1, world)]
    targets = [buf.open_peer(handles[p]) + 8 * n * rank for p in peers]
    views = [torch.as_tensor(_CudaView(t, n, "<f8"), device=cu) for t in targets]

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], device=cu)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sent = 8 * n * (world - 1)

    def report(name, ms):
        if rank == 0:
            print(f"{world} GPUs, {mib} MiB per slice: {name}: {ms:.3f} ms = {sent / (ms * 1e-3) / 1e9:.0f} GB/s per direction per GPU", flush=True)

    def copies():
        for v in views:
            v.copy_(src, non_blocking=True)
    report("cudaMemcpyAsync per peer, one stream", timed(copies))
    if world > 2:
        streams = [torch.cuda.Stream(priority=-1) for _ in views]
        done = [torch.cuda.Event() for _ in views]

        def copies_fanned():
            main = torch.cuda.current_stream()
            start = torch.cuda.Event()
            start.record(main)
            for v, st, ev in zip(views, streams, done):
                st.wait_event(start)
                with torch.cuda.stream(st):
                    v.copy_(src, non_blocking=True)
                ev.record(st)
                main.wait_event(ev)
        report(f"cudaMemcpyAsync per peer, {len(views)} streams", timed(copies_fanned))
    quick = world > 2      # multi-GPU time is charged per GPU: the short list
    for width in (32, 16):
        os.environ["SPMV_B200_PUSH_WIDTH"] = str(width)
        for unroll in ((2,) if quick else (1, 2, 4)):
            os.environ["SPMV_B200_PUSH_UNROLL"] = str(unroll)
            for ctas in ((148, 296, 592) if width == 32 or not quick else (296,)):
                report(f"vec_push {width}-byte lanes, unroll {unroll}, ctas {ctas}", timed(lambda: device.vec_push(src, n, targets, ctas=ctas)))
    # what landed
    dist.barrier()
    torch.cuda.synchronize()
    whole = buf.as_tensor()
    ok = all(bool((whole[p * n: (p + 1) * n] == float(p + 1)).all()) for p in range(world))
    flag = torch.tensor([1 if ok else 0], device=cu)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("PUSH_PROBE OK" if int(flag.item()) else "PUSH_PROBE FAILED", flush=True)
    dist.barrier()
    buf.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
