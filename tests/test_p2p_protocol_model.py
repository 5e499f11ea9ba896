"""Model check of the peer-memory all-reduce protocol (pointcloud_rl_b200/csrc/p2p.cu) on the CPU.

The kernel's flag protocol -- "ready" signal, wait for everybody's, reduce the own slice in place into every rank's
buffer, "done" signal, wait for everybody's, per-channel epochs kept locally -- is restated as one coroutine per rank and
run under thousands of random interleavings (every shared-memory access is a scheduling point).  Between two reductions a
rank overwrites its buffer with fresh gradients the moment ITS kernel has returned, exactly what the next update does.
Checked: no interleaving deadlocks, every rank ends every reduction with the exact sum of that epoch's inputs (so nobody
read a peer's buffer after the peer had moved on, and nobody's in-place store clobbered a slice somebody still had to read),
two channels in flight do not see each other's flags.  It pins the protocol's logic, not CUDA's memory model."""
import random

import numpy as np
import pytest

MAX_WORLD = 16


def flag_index(channel, which, src):
    return (channel * 2 + which) * MAX_WORLD + src


class Rank:
    def __init__(self, r, world, n, channels):
        self.r, self.world = r, world
        self.buf = np.zeros(n, dtype=np.int64)  # integer "gradients": sums are exact, any stale read shows up
        self.flags = np.zeros(channels * 2 * MAX_WORLD, dtype=np.int64)
        self.epoch = [0] * channels  # local device memory, advanced by the kernel itself


def allreduce(me, ranks, channel, lo, hi):
    """One kernel launch on rank `me` (a generator: every yield is a point where another rank may run)."""
    world = me.world
    e = me.epoch[channel] + 1
    for p in range(world):  # 1. my gradients are final
        ranks[p].flags[flag_index(channel, 0, me.r)] = e
        yield
    for p in range(world):
        while me.flags[flag_index(channel, 0, p)] < e:
            yield
    n = hi - lo
    per = (n + world - 1) // world
    s_lo, s_hi = min(lo + per * me.r, hi), min(lo + per * (me.r + 1), hi)
    for i in range(s_lo, s_hi):  # 2. my slice: load from everybody, sum in rank order, store to everybody
        acc = 0
        for p in range(world):
            acc += int(ranks[p].buf[i])
            yield
        for p in range(world):
            ranks[p].buf[i] = acc
            yield
    for p in range(world):  # 3. my slice is written everywhere
        ranks[p].flags[flag_index(channel, 1, me.r)] = e
        yield
    for p in range(world):
        while me.flags[flag_index(channel, 1, p)] < e:
            yield
    me.epoch[channel] = e


def rank_program(me, ranks, plan, rng, log):
    """The stream of one rank: for every step of the plan, fresh gradients, then the reduction(s) of that step.  A step
    with two ranges runs them as two concurrently scheduled kernels on two channels (two streams)."""
    for step, ranges in enumerate(plan):
        for ch, (lo, hi) in enumerate(ranges):
            vals = rng.integers(-1000, 1000, size=hi - lo)
            me.buf[lo:hi] = vals  # the next update's gradients overwrite the range: legal as soon as MY kernel returned
            log[(step, ch, me.r)] = vals.copy()
            yield
        kernels = [allreduce(me, ranks, ch, lo, hi) for ch, (lo, hi) in enumerate(ranges)]
        while kernels:
            k = kernels[rng.integers(len(kernels))]
            try:
                next(k)
            except StopIteration:
                kernels.remove(k)
            yield
        for ch, (lo, hi) in enumerate(ranges):  # what follows on the stream (Adam) reads the complete sum
            log[("out", step, ch, me.r)] = me.buf[lo:hi].copy()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_protocol_has_no_deadlock_and_sums_exactly_under_random_interleavings(world):
    n = 37
    plan = [[(0, 20), (20, 37)], [(0, 20)], [(0, 20), (20, 37)], [(20, 37)], [(0, 20), (20, 37)]]
    for seed in range(60):
        rng = np.random.default_rng(seed)
        sched = random.Random(seed)
        ranks = [Rank(r, world, n, channels=2) for r in range(world)]
        log = {}
        progs = [rank_program(ranks[r], ranks, plan, np.random.default_rng(1000 * seed + r), log) for r in range(world)]
        live = list(range(world))
        # skewed scheduling: some seeds let one rank run far ahead before the others move at all
        weights = [sched.choice([1, 1, 1, 20]) for _ in range(world)]
        steps = 0
        while live:
            r = sched.choices(live, weights=[weights[x] for x in live])[0]
            try:
                next(progs[r])
            except StopIteration:
                live.remove(r)
            steps += 1
            assert steps < 5_000_000, f"deadlock or livelock (seed {seed}, world {world})"
        for step, ranges in enumerate(plan):
            for ch, (lo, hi) in enumerate(ranges):
                want = sum(log[(step, ch, r)] for r in range(world))
                for r in range(world):
                    assert np.array_equal(log[("out", step, ch, r)], want), (seed, step, ch, r)
        del rng
