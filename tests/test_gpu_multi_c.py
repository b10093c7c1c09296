"""The single-process multi-GPU entry points (spmv_b200_multi_*, include/spmv_b200.h) driven from PLAIN C
(tests/c/multi_check.c -> sparsematrixvectormultiplication_b200/multi_check): CSR and HLL, fused MAILBOX and NCCL ALLGATHER
exchange, on every GPU count the box offers (1 on the driver's test box; 2+ under gpurun --gpus N).  The C program checks
N GPUs against one GPU; this wrapper pins the single-GPU lambda to the CPU oracle's power iteration."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
EXE = ROOT / "sparsematrixvectormultiplication_b200" / "multi_check"


def _run(ngpus, n):
    out = subprocess.run([str(EXE), str(ngpus), str(n)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "MULTI_CHECK OK" in out.stdout, out.stdout[-4000:] + out.stderr[-2000:]
    return [float(m.group(1)) for m in re.finditer(r"^lambda .* iters 12: (\S+) ", out.stdout, re.M)]


def test_c_entry_points_match_the_oracle(port):
    import torch
    from sparsematrixvectormultiplication_b200 import synth
    assert EXE.exists(), "multi_check was not built (make -C sparsematrixvectormultiplication_b200/csrc)"
    n = 24
    rp, ci, va = synth.lap3d_csr(n)
    _, _, lam_ref = port.power_iteration(rp, ci, va, np.ones(n ** 3), 12)
    counts = sorted({1, min(2, torch.cuda.device_count()), min(4, torch.cuda.device_count()), torch.cuda.device_count()})
    for ngpus in counts:
        lams = _run(ngpus, n)
        assert len(lams) == 8            # 1-GPU reference, csr/hll x mailbox/allgather/allgather_peer, host CSR
        for lam in lams:
            assert abs(lam - lam_ref[-1]) <= 1e-12 * lam_ref[-1], (ngpus, lam, lam_ref[-1])
