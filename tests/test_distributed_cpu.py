"""Host-side logic of the row-partitioned iterated product, world_size 2 and 3 over gloo on the CPU:
the nnz-balanced partition, the exchange plan (allgather and halo) and the refresh routines.  The local
product is supplied by the CPU oracle here (tests may use it; the product never does) -- on the GPU the
same plan drives the CUDA kernels (tests/test_gpu_parity.py, bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sparsematrixvectormultiplication_b200 import host, partition, synth
from sparsematrixvectormultiplication_b200.distributed import (ExchangePlan, exchange_allgather, exchange_halo, gather_needs,
                                                               padded_layout, remap_columns_host, unpad)


def test_closed_form_partition_is_the_reference_greedy_rule(checker):
    for kind, p0, twin in ((synth.SYNTH_LAP2D, 23, synth.lap2d_csr), (synth.SYNTH_LAP3D, 9, synth.lap3d_csr)):
        rp, _, _ = twin(p0)
        for T in (1, 2, 3, 4, 8, 16):
            mine = partition.synth_partition(kind, p0, 0, 0, T)
            s, e = checker.partition_rows(rp, T)
            assert mine == [(int(a), int(b)) for a, b in zip(s, e)], (kind, T)
            assert mine == partition.partition_rows(rp, T)
    rp, _, _ = synth.uniform_csr(1000, 640, 32)
    assert partition.synth_partition(synth.SYNTH_UNIFORM, 1000, 640, 32, 8) == partition.partition_rows(rp, 8)


def test_survey_partition_shape_128_cubed():
    """128^3 7-point Laplacian over 8 parts -- end parts get ~0.9% more rows."""
    parts = partition.synth_partition(synth.SYNTH_LAP3D, 128, 0, 0, 8)
    assert [e - s for s, e in parts] == [263922, 261558, 261558, 261540, 261539, 261558, 261558, 263919]


def test_hack_aligned_cuts():
    parts = partition.hack_aligned([(0, 70), (70, 131), (131, 200)], 200)
    assert parts == [(0, 64), (64, 128), (128, 200)]


def test_exchange_plan_halo_of_a_3d_laplacian():
    n, world = 8, 4
    parts = partition.synth_partition(synth.SYNTH_LAP3D, n, 0, 0, world)
    needs = []
    for s, e in parts:
        rp, ci, va = synth.lap3d_csr(n, s, e)
        needs.append((int(ci.min()), int(ci.max()) + 1))
    for r in range(world):
        plan = ExchangePlan.build(parts, needs, r)
        assert {p for p, _, _ in plan.recvs} <= {r - 1, r + 1}      # only neighbours
        assert plan.halo_doubles_received() <= 2 * n * n            # one plane from each side
        assert plan.allgather_doubles_received() == n ** 3 - (parts[r][1] - parts[r][0])
        for peer, lo, hi in plan.sends:                              # my sends are the peer's recvs
            assert (r, lo, hi) in ExchangePlan.build(parts, needs, peer).recvs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, kind, p0, iters, mode, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        chk = O.Restated()
        twin = synth.lap3d_csr if kind == synth.SYNTH_LAP3D else synth.lap2d_csr
        M = p0 ** 3 if kind == synth.SYNTH_LAP3D else p0 * p0
        parts = partition.synth_partition(kind, p0, 0, 0, world)
        lo, hi = parts[rank]
        rp, ci, va = twin(p0, lo, hi)
        need = (int(ci.min()), int(ci.max()) + 1)
        plan = ExchangePlan.build(parts, gather_needs(need, world, "cpu"), rank)
        if mode == "padded":   # AllgatherPowerIteration's host logic: padded rank-major x, remapped columns, ONE in-place all-gather
            starts, stride = padded_layout(parts)
            ci_pad = remap_columns_host(ci, starts, stride)
            xg = torch.ones(world * stride, dtype=torch.float64)
            own = xg[rank * stride: (rank + 1) * stride]
            for _ in range(iters):
                dist.all_gather_into_tensor(xg, own)           # refresh first, as the GPU class does
                y = torch.from_numpy(chk.spmv_csr_serial(rp, ci_pad, va, xg.numpy()))
                ss = (y * y).sum().reshape(1)
                dist.all_reduce(ss)
                own[:hi - lo] = y / ss.sqrt()
            dist.all_gather_into_tensor(xg, own)
            np.save(os.path.join(out_dir, f"x_{rank}.npy"), unpad(xg, parts, stride).numpy())
            np.save(os.path.join(out_dir, f"need_{rank}.npy"), np.array(need))
            return
        x = torch.ones(M, dtype=torch.float64)
        for _ in range(iters):
            y = torch.from_numpy(chk.spmv_csr_serial(rp, ci, va, x.numpy()))
            ss = (y * y).sum().reshape(1)
            dist.all_reduce(ss)
            x[lo:hi] = y / ss.sqrt()
            (exchange_allgather if mode == "allgather" else exchange_halo)(x, plan)
        np.save(os.path.join(out_dir, f"x_{rank}.npy"), x.numpy())
        np.save(os.path.join(out_dir, f"need_{rank}.npy"), np.array(need))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(2, "allgather"), (2, "halo"), (3, "halo"), (2, "padded"), (3, "padded")])
def test_power_iteration_two_and_three_ranks_gloo(tmp_path, port, world, mode):
    kind, p0, iters = synth.SYNTH_LAP3D, 7, 6
    mp.spawn(_worker, args=(world, _free_port(), kind, p0, iters, mode, str(tmp_path)), nprocs=world, join=True)
    rp, ci, va = synth.lap3d_csr(p0)
    x_ref, _, _ = port.power_iteration(rp, ci, va, np.ones(p0 ** 3), iters)
    parts = partition.synth_partition(kind, p0, 0, 0, world)
    for r in range(world):
        x = np.load(tmp_path / f"x_{r}.npy")
        lo, hi = np.load(tmp_path / f"need_{r}.npy") if mode == "halo" else (0, p0 ** 3)   # allgather / padded: whole vector
        lo, hi = min(lo, parts[r][0]), max(hi, parts[r][1])
        assert np.max(np.abs(x[lo:hi] - x_ref[lo:hi])) <= 1e-12 * np.max(np.abs(x_ref)), (r, mode)
