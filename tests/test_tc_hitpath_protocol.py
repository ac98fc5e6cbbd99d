"""Hand-over protocols added to the tcgen05 filter in round 2 (csrc/knn_filter_tc.cu), simulated on the CPU under a
random scheduler (sequentially consistent memory: what the kernel's fences are there to provide).

1. The multi-tile threshold bootstrap makes every role walk a sequence of ``boot + nt`` tiles -- tiles ``0 .. boot-1``
   first, for the threshold alone, then ``0 .. nt-1``: the TMA producer and the two MMA issuers by one running index,
   the sixteen epilogue warps in two loops that share the parity and ring-slot counters.  Checked: no deadlock, every
   epilogue warp reads the accumulator of the tile it expects in both loops, ring slots and accumulator buffers are
   never overwritten early, no waiter falls two phases behind.
2. The slab hit path: an epilogue lane waits for its slab (header word 0), writes the 32 scores, then the header, then
   raises its bit in its warp's pending word with a reduction; the selection warp snapshots the pending words, takes
   the bits it saw off each word with an AND-reduction, reads header and slab, hands the slab back (header 0) and
   stops when every epilogue warp has reported and nothing is pending.  Checked: every hit is consumed exactly once
   with the scores its owner wrote, none is lost at the end, no deadlock.
"""
import random

import pytest

EPI_WARPS = 16


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.completed = count, count, 0, 0

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


def boot_tiles(nt, cap=128, div=12):
    """The kernel's rule: a twelfth of the unit (a sixth below 128 tiles), even, at most `cap`, none under 4."""
    boot = (nt // (div // 2 if nt < 128 else div)) & ~1
    return cap if boot > cap else (0 if boot < 4 else boot)


class PipelineSim:
    """Producer, two MMA issuers (even / odd tiles of the running index) and sixteen epilogue warps."""

    def __init__(self, ns, nt, boot, seed):
        self.ns, self.nt, self.boot, self.ntv = ns, nt, boot, nt + boot
        self.rng = random.Random(seed)
        self.full = [MBar(1) for _ in range(ns)]
        self.free = [MBar(EPI_WARPS) for _ in range(ns)]
        self.acc_full = [MBar(1), MBar(1)]
        self.acc_free = [MBar(EPI_WARPS), MBar(EPI_WARPS)]
        self.slot_tile = [None] * ns
        self.acc_tile = [None, None]
        self.tma_q, self.mma_q = [], []
        self.read = [[] for _ in range(EPI_WARPS)]

    def tile_of(self, n):
        return n if n < self.boot else n - self.boot

    def wait(self, bar, parity, phase_index):
        while not bar.try_wait(parity):
            yield
        assert bar.completed >= phase_index + 1, 'woke before the intended phase'
        assert bar.completed <= phase_index + 2, 'barrier ran two phases ahead of a waiter'

    def producer(self):
        slot, rnd = 0, 0
        for n in range(self.ntv):
            if rnd > 0:
                yield from self.wait(self.free[slot], (rnd - 1) & 1, rnd - 1)
            bar = self.full[slot]
            bar.arrive_expect_tx(1)

            def land(slot=slot, n=n, bar=bar):
                self.slot_tile[slot] = (n, self.tile_of(n))
                bar.complete_tx(1)
            self.tma_q.append(land)
            yield
            slot += 1
            if slot == self.ns:
                slot, rnd = 0, rnd + 1

    def mma(self, first):
        for it in range(first, self.ntv, 2):
            buf, slot, rnd = it & 1, it % self.ns, it // self.ns
            yield from self.wait(self.full[slot], rnd & 1, rnd)
            assert self.slot_tile[slot] == (it, self.tile_of(it)), 'MMA read a ring slot holding another tile'
            if it >= 2:
                yield from self.wait(self.acc_free[buf], ((it >> 1) - 1) & 1, (it >> 1) - 1)

            def finish(slot=slot, it=it, buf=buf):
                assert self.slot_tile[slot][0] == it, 'ring slot overwritten while the MMA was reading it'
                self.acc_tile[buf] = (it, self.tile_of(it))
                self.acc_full[buf].arrive()
            self.mma_q.append(finish)
            yield

    def epilogue(self, w):
        state = {'parity': 0, 'slot': 0, 'v': 0}     # shared by the two loops, as in the kernel

        def one(buf, expect_tile):
            v = state['v']
            yield from self.wait(self.acc_full[buf], state['parity'], v >> 1)
            assert self.acc_tile[buf] == (v, expect_tile), 'accumulator holds another tile'
            yield                                     # the TMEM load
            assert self.acc_tile[buf] == (v, expect_tile), 'accumulator overwritten before it was read'
            self.read[w].append(expect_tile)
            self.acc_free[buf].arrive()
            self.free[state['slot']].arrive()
            state['slot'] = (state['slot'] + 1) % self.ns
            state['v'] = v + 1
            for _ in range(self.rng.randrange(3)):    # max tree, sometimes a hit
                yield

        for j in range(0, self.boot, 2):              # bootstrap loop: boot is even
            yield from one(0, j)
            yield from one(1, j + 1)
            state['parity'] ^= 1
        for it in range(0, self.nt, 2):               # tile loop
            yield from one(0, it)
            if it + 1 < self.nt:
                yield from one(1, it + 1)
            state['parity'] ^= 1

    def run(self):
        agents = [self.producer(), self.mma(0), self.mma(1)] + [self.epilogue(w) for w in range(EPI_WARPS)]
        live = list(range(len(agents)))
        idle = 0
        while live:
            queues = [q for q in (self.tma_q, self.mma_q) if q]
            if queues and self.rng.random() < 0.5:
                self.rng.choice(queues).pop(0)()      # asynchronous completions: in issue order per engine
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
            assert idle < 50000, 'deadlock: no progress'
        want = list(range(self.boot)) + list(range(self.nt))
        for w in range(EPI_WARPS):
            assert self.read[w] == want

    def snapshot(self):
        bars = self.full + self.free + self.acc_full + self.acc_free
        return tuple((b.completed, b.pending, b.tx) for b in bars)


@pytest.mark.parametrize('ns', [2, 3, 7, 8])
@pytest.mark.parametrize('nt', [1, 2, 5, 24, 25, 49, 131])
def test_bootstrap_tile_sequence(ns, nt):
    boot = boot_tiles(nt)
    assert boot % 2 == 0 and boot <= nt
    for seed in range(4):
        PipelineSim(ns, nt, boot, seed).run()


def test_boot_rule():
    assert [boot_tiles(n) for n in (0, 23, 24, 49, 127, 128, 360, 1535, 1536, 7813, 125000)] == \
        [0, 0, 4, 8, 20, 10, 30, 126, 128, 128, 128]


class SlabSim:
    """`warps` epilogue warps of `lanes` lanes each, one selection warp."""

    def __init__(self, warps, lanes, tiles, hit_rate, seed):
        self.rng = random.Random(seed)
        self.warps, self.lanes, self.tiles, self.hit_rate = warps, lanes, tiles, hit_rate
        n = warps * lanes
        self.hdr = [0] * n
        self.slab = [[None] * 4 for _ in range(n)]     # four words stand for the eight 16-byte stores
        self.pend = [0] * warps
        self.done = 0
        self.sent, self.got = [], []

    def lane(self, w, l):
        t = w * self.lanes + l
        for tile in range(self.tiles):
            for g in range(2):
                if self.rng.random() >= self.hit_rate:
                    continue
                while self.hdr[t] != 0:                # the slab is still taken
                    yield
                ev = (t, tile, g)
                for i in range(4):
                    self.slab[t][i] = (ev, i)
                    yield
                self.hdr[t] = 0x80000000 | (tile * 128 + 32 * g)
                yield                                  # fence
                self.pend[w] |= 1 << l                 # red.shared.or
                self.sent.append(ev)
                yield
        return

    def warp_done(self, lanes_left):
        # every lane of the warp has pushed what it had: the warp reports (atomic add on the done counter)
        while lanes_left[0] > 0:
            yield
        self.done += 1

    def selection(self):
        while True:
            done = self.done                           # read BEFORE the pending words
            yield
            snap = list(self.pend)
            yield
            progress = False
            for w in range(self.warps):
                if snap[w]:
                    self.pend[w] &= ~snap[w]           # red.shared.and with the bits seen
                    yield
            for w in range(self.warps):
                bits = snap[w]
                l = 0
                while bits:
                    if bits & 1:
                        t = w * self.lanes + l
                        hdr = self.hdr[t]
                        assert hdr & 0x80000000, 'pending bit without a published slab'
                        yield
                        words = list(self.slab[t])
                        ev = words[0][0]
                        assert all(wd == (ev, i) for i, wd in enumerate(words)), 'slab read while its owner was writing'
                        assert (hdr & 0x7fffffff) == ev[1] * 128 + 32 * ev[2]
                        yield
                        self.hdr[t] = 0                # slab back to its owner
                        self.got.append(ev)
                        progress = True
                    bits >>= 1
                    l += 1
            if not progress and done == self.warps:
                return
            yield

    def run(self):
        agents, left = [], []
        for w in range(self.warps):
            cnt = [self.lanes]
            left.append(cnt)

            def wrap(gen, cnt=cnt):
                yield from gen
                cnt[0] -= 1
            agents += [wrap(self.lane(w, l)) for l in range(self.lanes)]
            agents.append(self.warp_done(cnt))
        agents.append(self.selection())
        live = list(range(len(agents)))
        steps = 0
        while live:
            i = self.rng.choice(live)
            try:
                next(agents[i])
            except StopIteration:
                live.remove(i)
            steps += 1
            assert steps < 5_000_000, 'deadlock: no progress'
        assert sorted(self.got) == sorted(self.sent), 'a hit was lost or consumed twice'
        assert all(h == 0 for h in self.hdr) and all(p == 0 for p in self.pend)


@pytest.mark.parametrize('hit_rate', [0.02, 0.3, 1.0])
def test_slab_hand_over(hit_rate):
    for seed in range(6):
        SlabSim(warps=4, lanes=4, tiles=12, hit_rate=hit_rate, seed=seed).run()
