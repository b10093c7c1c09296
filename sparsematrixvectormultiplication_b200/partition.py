"""nnz-balanced contiguous row partition over GPUs.

It is the reference's own thread partitioner (prepare_thread_distribution, reference
src/csr_matrix.c:167-266) with GPUs in the role of threads: walk the rows, close a part as soon as
the nonzeros gathered since the previous cut reach ceil(total / parts), give the rest to the last
part, drop parts without nonzeros.  Two forms:

* partition_rows(row_ptr, parts)           -- calls the C function on a host row_ptr;
* partition_rows_by_offset(offset, M, parts) -- the same rule for matrices whose row offsets are
  a closed form (the synthetic Laplacians at 134 M rows): each cut is found by bisection on the
  monotone offset function instead of a linear walk.
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import numpy as np


def partition_rows(row_ptr, parts: int) -> List[Tuple[int, int]]:
    from . import host
    row_ptr = np.ascontiguousarray(row_ptr, np.int32)
    s, e = host.prepare_thread_distribution(len(row_ptr) - 1, row_ptr, parts, int(row_ptr[-1]))
    return [(int(a), int(b)) for a, b in zip(s, e)]


def partition_rows_by_offset(offset: Callable[[int], int], M: int, parts: int) -> List[Tuple[int, int]]:
    """offset(r) = number of nonzeros in rows [0, r), non-decreasing."""
    if M <= 0 or parts <= 0:
        return []
    parts = min(parts, M)
    total = offset(M)
    target = -(-total // parts)
    out: List[Tuple[int, int]] = []
    start = 0
    for part in range(parts):
        if start >= M:
            break
        if part == parts - 1:
            end = M
        else:
            base = offset(start)
            # smallest end in (start, M] with offset(end) - base >= target; M when never reached
            lo, hi = start + 1, M
            if offset(M) - base < target:
                end = M
            else:
                while lo < hi:
                    mid = (lo + hi) // 2
                    if offset(mid) - base >= target:
                        hi = mid
                    else:
                        lo = mid + 1
                end = lo
        if offset(end) - offset(start) > 0:
            out.append((start, end))
        start = end
    return out


def synth_partition(kind: int, p0: int, p1: int, p2: int, parts: int) -> List[Tuple[int, int]]:
    from . import synth
    M = {synth.SYNTH_LAP2D: p0 * p0, synth.SYNTH_LAP3D: p0 ** 3, synth.SYNTH_UNIFORM: p0}[kind]
    return partition_rows_by_offset(lambda r: synth.row_offset(kind, p0, p1, p2, r), M, parts)


def hack_aligned(parts: List[Tuple[int, int]], M: int, hack: int = 32) -> List[Tuple[int, int]]:
    """Move every interior cut down to a multiple of the hack size (HLL blocks are indivisible:
    the reference cuts HLL work on block boundaries, src/hll_matrix.c:471-498)."""
    cuts = [0] + [(e // hack) * hack for _, e in parts[:-1]] + [M]
    return [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
