"""Model check of the two in-kernel exchange protocols of the fused iterated product (no GPU needed).

The kernels (csrc/cuda/csr.cu: csr_row_fused_kernel + mail_publish / mail_wait_total, csr_row_async_kernel) order their
work with tags in peer mailboxes and reuse a small ring of buffers.  Whether a ring is deep enough is a question about
ALL interleavings the waits allow, not about the one a test run happens to produce, so the rules are restated here as a
discrete-event model and driven by random rank speeds:

  mailbox form  (spmv_b200_csr_spmv_fused_mail):  x in 2 buffers, sums in 2 parity slots per rank;
      launch k starts after the tags k of ALL ranks (their launch k-1 is finished) have arrived.
  async form    (spmv_b200_csr_spmv_fused_async): x in 3 buffers, sums in a 4-slot ring, one halo tag per neighbour;
      launch k starts after the sums of launch k-2 of ALL ranks and the halo tags >= k of its neighbours have arrived;
      the halo tag k+1 is raised when the boundary rows of launch k are stored (early), the sum at the end.

Checked for every launch of every rank: the halo it reads was written by the neighbour's previous launch, nobody writes
into a buffer a neighbour is still reading, every sum slot still carries the expected tag when it is read, and no launch
ever deadlocks.  Making a ring one slot shallower must make the model fail (the test of the test)."""
import heapq
import random

import pytest


class Violation(AssertionError):
    pass


def simulate(world, launches, seed, mode, x_ring, sum_ring, max_delay=0.05):
    """Event-driven run.  Returns the number of launches completed per rank (all == launches unless deadlocked)."""
    rng = random.Random(seed)
    neighbours = {r: [p for p in (r - 1, r + 1) if 0 <= p < world] for r in range(world)}
    speed = [rng.uniform(0.6, 1.6) for _ in range(world)]          # per-rank kernel duration scale
    # state
    halo_content = [[{p: -1 for p in neighbours[r]} for _ in range(x_ring)] for r in range(world)]  # launch that wrote it
    reading = [None] * world                                        # x buffer index rank r is gathering from right now
    sum_tag = [[[0] * world for _ in range(sum_ring)] for _ in range(world)]   # [owner][slot][sender] = tag
    halo_tag = [{p: 0 for p in neighbours[r]} for r in range(world)]            # async: latest launch+1 whose halo landed
    done = [0] * world                                              # launches finished
    started = [False] * world
    events = []                                                     # (time, seq, kind, rank, launch, peer)
    seq = [0]

    def push(t, kind, r, k, peer=None):
        seq[0] += 1
        heapq.heappush(events, (t, seq[0], kind, r, k, peer))

    def delay():  # time a tag / sum needs to become visible on another rank, in units of one launch
        return rng.uniform(0.001, max_delay)

    def can_start(r, k):
        if mode == "mailbox":
            if k == 0:
                return True
            slot = (k - 1) % sum_ring
            return all(sum_tag[r][slot][s] == k for s in range(world))
        if k >= 2 and not all(sum_tag[r][(k - 2) % sum_ring][s] == k - 1 for s in range(world)):
            return False
        return k == 0 or all(halo_tag[r][p] >= k for p in neighbours[r])

    def try_start(r, now):
        k = done[r]
        if started[r] or k >= launches or not can_start(r, k):
            return
        started[r] = True
        cur = k % x_ring
        # (1) the halo this launch gathers must be what the neighbours' launch k-1 stored
        if k > 0:
            for p in neighbours[r]:
                if halo_content[r][cur][p] != k - 1:
                    raise Violation(f"{mode}: rank {r} launch {k} reads halo of rank {p} written by launch {halo_content[r][cur][p]}")
        # sums it consumes: still the expected tags (checked in can_start), nothing else to do
        reading[r] = cur
        dur = speed[r] * rng.uniform(0.9, 1.1)
        boundary_at = now + (0.08 * dur if mode == "async" else dur * rng.uniform(0.0, 1.0))
        push(boundary_at, "boundary", r, k)
        push(now + dur, "end", r, k)

    for r in range(world):
        try_start(r, 0.0)
    now = 0.0
    while events:
        now, _, kind, r, k, peer = heapq.heappop(events)
        if kind == "boundary":
            nxt = (k + 1) % x_ring
            for p in neighbours[r]:
                # (2) writing into the neighbour's next buffer: it must not be gathering from that buffer right now
                if reading[p] == nxt:
                    raise Violation(f"{mode}: rank {r} launch {k} writes rank {p}'s buffer {nxt} while it is being read")
                halo_content[p][nxt][r] = k
                if mode == "async":
                    push(now + delay(), "halo_tag", p, k, r)
        elif kind == "halo_tag":
            halo_tag[r][peer] = max(halo_tag[r][peer], k + 1)
            try_start(r, now)
        elif kind == "end":
            reading[r] = None
            done[r] = k + 1
            started[r] = False
            for dest in range(world):
                push(now + (delay() if dest != r else 0.0), "sum", dest, k, r)
            try_start(r, now)
        elif kind == "sum":
            slot = k % sum_ring
            # (3) overwriting a sum slot: every reader of the previous occupant must be past it.  Reader of tag t in this
            # slot is launch t (mailbox) / t + 1 (async) of rank r; it is past it once that launch has STARTED.
            old = sum_tag[r][slot][peer]
            if old:
                reader_launch = old if mode == "mailbox" else old + 1
                if reader_launch < launches and (done[r] < reader_launch or (done[r] == reader_launch and not started[r])):
                    raise Violation(f"{mode}: sum slot {slot} of rank {r} (from rank {peer}, tag {old}) overwritten before launch {reader_launch} read it")
            sum_tag[r][slot][peer] = k + 1
            try_start(r, now)
    return done


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("mode,x_ring,sum_ring", [("mailbox", 2, 2), ("async", 3, 4)])
def test_protocol_is_hazard_free_under_random_schedules(world, mode, x_ring, sum_ring):
    for seed in range(60):
        done = simulate(world, 25, seed, mode, x_ring, sum_ring)
        assert done == [25] * world, f"deadlock: {done}"


@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("mode,x_ring,sum_ring", [("mailbox", 2, 2), ("async", 3, 4)])
def test_protocol_survives_slow_messages(world, mode, x_ring, sum_ring):
    """Tags that take LONGER than a whole launch to arrive (far beyond NVLink reality) slow everything down but break
    nothing: the ring depths do not rely on timing."""
    for seed in range(40):
        done = simulate(world, 20, seed, mode, x_ring, sum_ring, max_delay=3.0)
        assert done == [20] * world, f"deadlock: {done}"


def _violations(mode, x_ring, sum_ring, max_delay):
    caught = 0
    for seed in range(200):
        try:
            simulate(4, 25, seed, mode, x_ring, sum_ring, max_delay=max_delay)
        except Violation:
            caught += 1
    return caught


def test_shallower_rings_are_caught():
    """The test of the test: with two x buffers the async form overwrites a halo that is still being read, with two sum
    slots a sum before it was consumed; the depths the kernels use survive even messages slower than launches."""
    assert _violations("async", 2, 4, 0.05) > 0
    assert _violations("async", 3, 2, 0.05) > 0
    assert _violations("async", 3, 4, 3.0) == 0
    assert _violations("mailbox", 2, 2, 3.0) == 0
