"""Barrier protocol of the tcgen05 filter engine (csrc/knn_filter_tc.cu), simulated on the CPU.

The kernel has three roles that only meet at mbarriers: the TMA producer, the MMA issuer and eight epilogue
warps.  This model restates their loops with the kernel's own slot / parity arithmetic and runs them under a
random scheduler with mbarrier semantics (phase completes when the pending arrivals and transaction bytes
reach zero; ``try_wait.parity(P)`` succeeds when the phase of parity P has completed).  Checked: no deadlock,
every reader sees the tile it expects (ring slots and accumulator buffers are never overwritten early), and no
waiter can fall two phases behind (which parity waits cannot express).
"""
import random

import pytest

EPI_WARPS = 8


class MBar:
    def __init__(self, count):
        self.count = count
        self.pending = count
        self.tx = 0
        self.completed = 0

    def _maybe_complete(self):
        if self.pending == 0 and self.tx == 0:
            self.completed += 1
            self.pending = self.count

    def arrive(self):
        assert self.pending > 0
        self.pending -= 1
        self._maybe_complete()

    def arrive_expect_tx(self, nbytes):
        self.tx += nbytes
        self.arrive()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._maybe_complete()

    def try_wait(self, parity):
        return (self.completed & 1) != parity


class Sim:
    def __init__(self, ns, nt, seed):
        self.ns, self.nt = ns, nt
        self.rng = random.Random(seed)
        self.full = [MBar(1) for _ in range(ns)]
        self.free = [MBar(EPI_WARPS) for _ in range(ns)]
        self.acc_full = [[MBar(1), MBar(1)], [MBar(1), MBar(1)]]   # [buffer][query block]
        self.acc_free = [[MBar(EPI_WARPS // 2), MBar(EPI_WARPS // 2)] for _ in range(2)]   # [buffer][query block]
        self.slot_tile = [None] * ns        # tile whose bytes are in the ring slot
        self.acc_tile = [[None, None], [None, None]]   # tile whose scores are in the accumulator buffer, per query block
        self.tma_q, self.mma_q = [], []     # asynchronous completions, in issue order per engine

    # -- a wait that also checks the waiter is not two phases late -------------------------------
    def wait(self, bar, parity, phase_index):
        while not bar.try_wait(parity):
            yield
        # the phase we meant must be complete, and the barrier may be at most one phase further
        assert bar.completed >= phase_index + 1, 'woke before the intended phase'
        assert bar.completed <= phase_index + 2, 'barrier ran two phases ahead of a waiter'

    def producer(self):
        slot, rnd = 0, 0
        for n in range(self.nt):
            if rnd > 0:
                yield from self.wait(self.free[slot], (rnd - 1) & 1, rnd - 1)
            bar = self.full[slot]
            bar.arrive_expect_tx(1)

            def land(slot=slot, n=n, bar=bar):
                self.slot_tile[slot] = n
                bar.complete_tx(1)
            self.tma_q.append(land)
            yield
            slot += 1
            if slot == self.ns:
                slot, rnd = 0, rnd + 1

    def mma(self):
        slot, rnd = 0, 0
        for it in range(self.nt):
            buf = it & 1
            yield from self.wait(self.full[slot], rnd & 1, rnd)
            assert self.slot_tile[slot] == it, 'MMA read a ring slot holding another tile'
            for mb in range(2):
                if it >= 2:
                    yield from self.wait(self.acc_free[buf][mb], ((it >> 1) - 1) & 1, (it >> 1) - 1)

                def finish(slot=slot, it=it, buf=buf, mb=mb):
                    assert self.slot_tile[slot] == it, 'ring slot overwritten while the MMA was reading it'
                    self.acc_tile[buf][mb] = it
                    self.acc_full[buf][mb].arrive()  # tcgen05.commit after the MMAs of this query block
                self.mma_q.append(finish)
                yield
            slot += 1
            if slot == self.ns:
                slot, rnd = 0, rnd + 1

    def epilogue(self, w):
        slot, rnd = 0, 0
        mb = w >> 2
        for it in range(self.nt):
            buf = it & 1
            yield from self.wait(self.acc_full[buf][mb], (it >> 1) & 1, it >> 1)
            for _ in range(4):                       # four tcgen05.ld groups
                assert self.acc_tile[buf][mb] == it, 'accumulator buffer overwritten before it was read'
                yield
            self.acc_free[buf][mb].arrive()
            # slow path of the last group: reads the operand tile (after observing its TMA barrier)
            yield from self.wait(self.full[slot], rnd & 1, rnd)
            assert self.slot_tile[slot] == it, 'slow path read a ring slot holding another tile'
            yield
            self.free[slot].arrive()
            slot += 1
            if slot == self.ns:
                slot, rnd = 0, rnd + 1

    def run(self):
        agents = [self.producer(), self.mma()] + [self.epilogue(w) for w in range(EPI_WARPS)]
        live = list(range(len(agents)))
        idle = 0
        while live:
            # asynchronous completions happen in issue order per engine but at arbitrary times
            queues = [q for q in (self.tma_q, self.mma_q) if q]
            if queues and self.rng.random() < 0.5:
                self.rng.choice(queues).pop(0)()
                idle = 0
                continue
            i = self.rng.choice(live)
            before = self.snapshot()
            try:
                next(agents[i])
            except StopIteration:
                live.remove(i)
                idle = 0
                continue
            idle = idle + 1 if self.snapshot() == before and not (self.tma_q or self.mma_q) else 0
            assert idle < 20000, 'deadlock: no progress'
        for q in (self.tma_q, self.mma_q):
            while q:
                q.pop(0)()

    def snapshot(self):
        bars = self.full + self.free + self.acc_full[0] + self.acc_full[1] + self.acc_free[0] + self.acc_free[1]
        return tuple((b.completed, b.pending, b.tx) for b in bars)


@pytest.mark.parametrize('ns', [2, 3, 4, 8])
@pytest.mark.parametrize('nt', [0, 1, 2, 3, 7, 33])
def test_tc_barrier_protocol(ns, nt):
    for seed in range(12):
        Sim(ns, nt, seed).run()
