#!/usr/bin/env python
"""Multi-GPU parity check, launched with torch.distributed.run (one rank per GPU):
the row-partitioned power iteration (halo and allgather refresh) must reproduce the single-GPU iteration
on every rank's owned + referenced part of x.  Rows of the 7-point Laplacian are summed by one lane in
index order on every rank, so the comparison is BITWISE apart from the norm (an N-rank all-reduce of
|y|^2 associates differently from the single-GPU sum): x is compared with 1e-13 relative tolerance and
lambda with 1e-13.  Prints 'DIST_CHECK OK' on rank 0."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from sparsematrixvectormultiplication_b200 import synth  # noqa: E402
from sparsematrixvectormultiplication_b200.distributed import (AllgatherPowerIteration, AsyncPowerIteration, FusedPowerIteration,  # noqa: E402
                                                               PeerAllgatherPowerIteration, PowerIteration)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, iters = int(sys.argv[1]) if len(sys.argv) > 1 else 48, 12
    ref = PowerIteration(synth.SYNTH_LAP3D, n, single=True)
    lam_ref = []
    for _ in range(iters):
        ref.step()
        lam_ref.append(ref.eigenvalue_estimate())
    ok = True
    for mode in ("halo", "allgather"):
        P = PowerIteration(synth.SYNTH_LAP3D, n, exchange=mode)
        lam = []
        for _ in range(iters):
            P.step()
            lam.append(P.eigenvalue_estimate())
        torch.cuda.synchronize()
        lo = min([P.row_begin] + [a for _, a, _ in P.plan.recvs]) if mode == "halo" else 0
        hi = max([P.row_end] + [b for _, _, b in P.plan.recvs]) if mode == "halo" else P.N
        err = float((P.x[lo:hi] - ref.x[lo:hi]).abs().max() / ref.x.abs().max())
        lam_err = max(abs(a - b) / b for a, b in zip(lam, lam_ref))
        # the product itself must be bitwise partition independent: same x in, same y out on the owned rows
        P.x.copy_(ref.x)
        P.A.spmv(P.x, P.y)
        ref.A.spmv(ref.x, ref.y)
        same = bool(torch.equal(P.y[:P.rows], ref.y[P.row_begin:P.row_end]))
        good = err <= 1e-13 and lam_err <= 1e-13 and same
        print(f"rank {rank}/{world} {mode}: rows [{P.row_begin},{P.row_end}) recv {P.plan.halo_doubles_received() if mode == 'halo' else P.plan.allgather_doubles_received()} "
              f"doubles/iter, x err {err:.2e}, lambda err {lam_err:.2e}, product bitwise {same} -> {'ok' if good else 'FAIL'}", flush=True)
        ok = ok and good
    for peer_stores, mailbox in ((True, False), (False, False), (True, True)):
        F = FusedPowerIteration(synth.SYNTH_LAP3D, n, peer_stores=peer_stores, mailbox=mailbox)
        lam = []
        for _ in range(iters):
            F.step()
            lam.append(F.eigenvalue_estimate())
        torch.cuda.synchronize()
        v = F.normalized_x()
        lo = min([F.row_begin] + [a for _, a, _ in F.plan.recvs])
        hi = max([F.row_end] + [b for _, _, b in F.plan.recvs])
        err = float((v[lo:hi] - ref.x[lo:hi]).abs().max() / ref.x.abs().max())
        lam_err = max(abs(a - b) / b for a, b in zip(lam, lam_ref))
        good = err <= 1e-12 and lam_err <= 1e-12
        print(f"rank {rank}/{world} fused peer_stores={peer_stores} mailbox={mailbox}: x err {err:.2e}, lambda err {lam_err:.2e} -> {'ok' if good else 'FAIL'}", flush=True)
        ok = ok and good
        F.close()
    # round 2: the two-launch form (CSR and HLL), the one-collective NCCL all-gather and the all-gather written against
    # peer memory (kernel push and copy-engine push); whole replica compared for the all-gather modes
    makers = (("split", lambda: FusedPowerIteration(synth.SYNTH_LAP3D, n, split=True), False),
              ("split hll", lambda: FusedPowerIteration(synth.SYNTH_LAP3D, n, split=True, fmt="hll"), False),
              ("allgather (one ncclAllGather)", lambda: AllgatherPowerIteration(synth.SYNTH_LAP3D, n), True),
              ("allgather_peer (copy engines)", lambda: PeerAllgatherPowerIteration(synth.SYNTH_LAP3D, n), True),
              ("allgather_peer (copy engines, 3 streams)", lambda: PeerAllgatherPowerIteration(synth.SYNTH_LAP3D, n, copy_streams=3), True),
              ("allgather_peer (copy engines + push kernel for the last peer)", lambda: PeerAllgatherPowerIteration(synth.SYNTH_LAP3D, n, kernel_peers=1), True),
              ("allgather_peer (push kernel)", lambda: PeerAllgatherPowerIteration(synth.SYNTH_LAP3D, n, copy_engine=False), True))
    for name, make, whole in makers:
        F = make()
        lam = []
        for _ in range(iters):
            F.step()
            lam.append(F.eigenvalue_estimate())
        v = F.normalized_x()
        lo = 0 if whole else min([F.row_begin] + [a for _, a, _ in F.plan.recvs])
        hi = F.N if whole else max([F.row_end] + [b for _, _, b in F.plan.recvs])
        err = float((v[lo:hi] - ref.x[lo:hi]).abs().max() / ref.x.abs().max())
        lam_err = max(abs(a - b) / b for a, b in zip(lam, lam_ref))
        good = err <= 1e-12 and lam_err <= 1e-12
        print(f"rank {rank}/{world} {name}: rows [{F.row_begin},{F.row_end}), x err {err:.2e} on [{lo},{hi}), lambda err {lam_err:.2e} -> {'ok' if good else 'FAIL'}", flush=True)
        ok = ok and good
        if hasattr(F, "close"):
            F.close()
    # the peer all-gather free-running on a small matrix (short launches, no host synchronisation): a slice or a sum read
    # before it landed, or a replica overwritten while a slow rank still reads it, would throw lambda off
    small = 24
    R = PowerIteration(synth.SYNTH_LAP3D, small, single=True)
    for _ in range(600):
        R.step()
    for engine in (True, False):
        S = PeerAllgatherPowerIteration(synth.SYNTH_LAP3D, small, copy_engine=engine)
        for _ in range(600):
            S.step()
        good = abs(S.eigenvalue_estimate() - R.eigenvalue_estimate()) / R.eigenvalue_estimate() <= 1e-12
        xerr = float((S.normalized_x() - R.x).abs().max() / R.x.abs().max())
        good = good and xerr <= 1e-11
        print(f"rank {rank}/{world} allgather_peer ({'copy engines' if engine else 'push kernel'}) stress {small}^3 x 600 iterations: lambda {S.eigenvalue_estimate():.15g} "
              f"vs {R.eigenvalue_estimate():.15g}, x err {xerr:.2e} -> {'ok' if good else 'FAIL'}", flush=True)
        ok = ok and good
        S.close()
    del R
    # asynchronous form: boundary rows first, halo tags, scale factor lagging one launch; twice (reset in between)
    G = AsyncPowerIteration(synth.SYNTH_LAP3D, n)
    for attempt in range(2):
        lam = []
        for it in range(iters):
            G.step()
            if it >= 1:
                lam.append(G.eigenvalue_estimate())
        v = G.normalized_x()
        lo = min([G.row_begin] + [a for _, a, _ in G.plan.recvs])
        hi = max([G.row_end] + [b for _, _, b in G.plan.recvs])
        err = float((v[lo:hi] - ref.x[lo:hi]).abs().max() / ref.x.abs().max())
        lam_err = max(abs(a - b) / b for a, b in zip(lam, lam_ref[1:]))
        good = err <= 1e-12 and lam_err <= 1e-12
        print(f"rank {rank}/{world} async attempt {attempt}: x err {err:.2e}, lambda err {lam_err:.2e} -> {'ok' if good else 'FAIL'}", flush=True)
        ok = ok and good
        G.reset(1.0)
    # and free-running (no host synchronisation between launches), which is how it is timed
    for _ in range(40):
        G.step()
    lam_free = G.eigenvalue_estimate()
    for _ in range(40 - iters):
        ref.step()
    good = abs(lam_free - ref.eigenvalue_estimate()) / ref.eigenvalue_estimate() <= 1e-12
    print(f"rank {rank}/{world} async free-running 40 launches: lambda {lam_free:.15g} vs {ref.eigenvalue_estimate():.15g} -> {'ok' if good else 'FAIL'}", flush=True)
    ok = ok and good
    G.close()
    # stress: short launches (small grid), many of them, no host synchronisation: a halo row or a sum that was read
    # before it landed would throw lambda off
    small = 24
    S = AsyncPowerIteration(synth.SYNTH_LAP3D, small)
    R = PowerIteration(synth.SYNTH_LAP3D, small, single=True)
    for _ in range(600):
        S.step()
    for _ in range(600):
        R.step()
    good = abs(S.eigenvalue_estimate() - R.eigenvalue_estimate()) / R.eigenvalue_estimate() <= 1e-12
    vs, lo, hi = S.normalized_x(), S.row_begin, S.row_end
    xerr = float((vs[lo:hi] - R.x[lo:hi]).abs().max() / R.x.abs().max())
    good = good and xerr <= 1e-11
    print(f"rank {rank}/{world} async stress {small}^3 x 600 launches: lambda {S.eigenvalue_estimate():.15g} vs {R.eigenvalue_estimate():.15g}, x err {xerr:.2e} -> {'ok' if good else 'FAIL'}", flush=True)
    ok = ok and good
    S.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_CHECK OK" if int(flag.item()) else "DIST_CHECK FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
