"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): tools/dist_check.py under torchrun."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_row_partitioned_power_iteration_matches_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", str(ROOT / "tools" / "dist_check.py"), "40"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DIST_CHECK OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
