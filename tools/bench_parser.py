#!/usr/bin/env python
"""read_matrix_market on a generated file: this repo's serial tokenizer and its parallel tokenizer.  CPU only.
Usage: python tools/bench_parser.py [entries]
(The reference's fscanf loop on the same file is timed by tests/test_host_api.py::test_parser_speed_report, because only
tests/, smoke() and bench.py may execute anything under oracle/.)"""
import os
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402

from sparsematrixvectormultiplication_b200 import host  # noqa: E402


def main():
    nz = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
    rng = np.random.default_rng(1)
    M = N = 2_000_000
    I = rng.integers(1, M + 1, nz)
    J = rng.integers(1, N + 1, nz)
    V = rng.standard_normal(nz)
    with tempfile.TemporaryDirectory() as d:
        path = Path(d) / "big.mtx"
        with open(path, "w") as f:
            f.write(f"%%MatrixMarket matrix coordinate real general\n{M} {N} {nz}\n")
            np.savetxt(f, np.column_stack([I, J, V]), fmt="%d %d %.17g")
        size = path.stat().st_size
        print(f"{nz} entries, {size/1e6:.0f} MB, {os.cpu_count()} logical cores")

        def timed(label, fn):
            t0 = time.perf_counter()
            out = fn()
            dt = time.perf_counter() - t0
            print(f"{label:42s} {dt:7.2f} s  {size/dt/1e6:8.1f} MB/s  {nz/dt/1e6:6.2f} M entries/s")
            return out
        os.environ["SPMV_B200_PARSER_PARALLEL_MIN_BYTES"] = str(1 << 60)
        a = timed("this repo, serial tokenizer", lambda: host.read_matrix_market(path))
        del os.environ["SPMV_B200_PARSER_PARALLEL_MIN_BYTES"]
        b = timed("this repo, parallel tokenizer (default)", lambda: host.read_matrix_market(path))
        assert np.array_equal(a.I, b.I) and np.array_equal(a.J, b.J) and np.array_equal(a.val.view(np.uint64), b.val.view(np.uint64))


if __name__ == "__main__":
    main()
