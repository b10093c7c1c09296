"""Shared fixtures.  `-m "not gpu"` runs here on the CPU; `-m gpu` runs on a B200.

Only the tests (and smoke()/bench.py's baseline legs) may touch oracle/ -- it is the checker.
"""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the native library and the oracle exist (cheap no-ops when up to date)."""
    import sparsematrixvectormultiplication_b200 as pkg
    try:
        import torch
        on_gpu_box = torch.cuda.is_available()
    except Exception:
        on_gpu_box = False
    # development container (no GPU): run make, a no-op when the library is newer than every source and a rebuild when
    # it is stale.  GPU box: the library that travelled with the snapshot is the one under test.
    if not on_gpu_box or not pkg.LIB_PATH.exists():
        pkg.build()
    from oracle import oracle as O
    if not (O.HERE / "liboracle.so").exists():
        O.build()


@pytest.fixture(scope="session")
def port():
    from oracle import oracle as O
    return O.Restated()


@pytest.fixture(scope="session")
def reference():
    from oracle import oracle as O
    if not O.reference_available():
        pytest.skip("oracle/_ref/libspmv_ref.so not built (needs /root/reference)")
    return O.Reference()


@pytest.fixture(scope="session")
def checker():
    """The strongest oracle available: the real reference when its .so exists, else the port."""
    from oracle import oracle as O
    return O.best_available()


def golden_names():
    return sorted(p.stem for p in GOLDEN.glob("*.npz"))


def load_golden(name):
    return np.load(GOLDEN / f"{name}.npz")


def ramp(n):
    """x_i = 1 + (i mod 7)/8 -- the second probe vector of tests/golden/make_golden.py."""
    return 1.0 + (np.arange(n) % 7) / 8.0


def random_coo(rng, M, N, nz, dup=True):
    from oracle import oracle as O
    I = rng.integers(0, M, nz).astype(np.int32)
    J = rng.integers(0, N, nz).astype(np.int32)
    if not dup and nz:
        key = np.unique(I.astype(np.int64) * N + J)
        rng.shuffle(key)
        I, J = (key // N).astype(np.int32), (key % N).astype(np.int32)
    return O.Coo(M, N, len(I), I, J, rng.standard_normal(len(I)))
