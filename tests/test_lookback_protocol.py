"""A CPU model of the two ordered-emission protocols of k_shade / k_capture (rpx_kernels.cuh:
tile_publish / tile_lookback and tile_publish_grouped / tile_lookback_grouped).

The CUDA code cannot run here, but the PROTOCOL can: tiles take tickets in order, publish their
aggregate at some later time, and resolve their exclusive prefix by reading state words that other tiles
write asynchronously.  The model executes the same decisions as the device code (same words, same flag /
count encodings, same "wait only for the entries in front of the nearest PREFIX" rule) under randomly
interleaved schedules (independent CTAs, and a resident wave moving in lock step) and checks that every
tile obtains the exact exclusive sum and that the run always terminates: a tile only ever waits on tiles
with lower tickets, all of which are resident or finished, so a persistent grid cannot deadlock.
(How many memory round trips a look-back costs is a latency question the model does not answer: that
is measured on the GPU, profiles/r01_notes.md.)
"""
import random

import pytest

FLAG_AGG, FLAG_PREFIX = 1 << 62, 2 << 62
VAL_MASK = (1 << 62) - 1
GROUP_ONE, GROUP_SUM_MASK = 1 << 56, (1 << 48) - 1


class Sim(object):
    def __init__(self, totals, grouped, resident, seed):
        self.totals, self.grouped = totals, grouped
        n = len(totals)
        self.state = [0] * n
        self.gagg = [0] * ((n + 31) // 32)
        self.gpre = [0] * ((n + 31) // 32)
        self.prefix = [None] * n
        self.rounds = [0] * n
        self.rng = random.Random(seed)
        self.resident = resident

    # ---- what thread 0 of a tile does right after the block scan
    def publish(self, t):
        self.state[t] = (FLAG_PREFIX if t == 0 else FLAG_AGG) | self.totals[t]
        if self.grouped:
            self.gagg[t >> 5] += GROUP_ONE | self.totals[t]

    # ---- one attempt of warp 0 to resolve the prefix; returns False if it has to keep polling
    def lookback(self, t):
        return self._grouped(t) if self.grouped else self._flat(t)

    def _window(self, words):
        """words: nearest first.  -> (ok, sum, found_prefix): usable once every word in front of the nearest
        PREFIX is there; the sum runs up to and including that PREFIX."""
        pref = next((i for i, w in enumerate(words) if (w >> 62) == 2), None)
        need = words if pref is None else words[:pref]
        if any((w >> 62) == 0 for w in need):
            return False, 0, False
        upto = words if pref is None else words[:pref + 1]
        return True, sum(w & VAL_MASK for w in upto), pref is not None

    def _flat(self, t):
        if t == 0:
            return self._done(t, 0)
        excl, t0, rounds = 0, t - 1, 0
        while True:
            words = [self.state[k] for k in range(t0, max(t0 - 32, -1), -1)]
            rounds += 1
            ok, s, found = self._window(words)
            if not ok:
                return False
            excl += s
            if found:
                self.rounds[t] = rounds
                return self._done(t, excl)
            t0 -= 32

    def _grouped(self, t):
        if t == 0:
            return self._done(t, 0)
        g, j = t >> 5, t & 31
        ok, excl, found = self._window([self.state[k] for k in range(t - 1, t - 1 - j, -1)])
        if not ok:
            return False
        rounds, gi = 1, g - 1
        while not found:
            words = []
            for k in range(gi, gi - 32, -1):
                if k < 0:
                    words.append(FLAG_PREFIX)          # before group 0: an empty PREFIX
                elif (self.gpre[k] >> 62) == 2:
                    words.append(self.gpre[k])
                elif (self.gagg[k] >> 56) == 32:
                    words.append(FLAG_AGG | (self.gagg[k] & GROUP_SUM_MASK))
                else:
                    words.append(0)
            ok, s, found = self._window(words)
            if not ok:
                return False
            excl += s
            gi -= 32
            rounds += 0 if gi == g - 33 else 1          # the first group window is loaded with the own-group one
        self.rounds[t] = rounds
        return self._done(t, excl)

    def _done(self, t, excl):
        self.prefix[t] = excl
        self.state[t] = FLAG_PREFIX | (excl + self.totals[t])
        if self.grouped and (t & 31) == 31:
            self.gpre[t >> 5] = FLAG_PREFIX | (excl + self.totals[t])
        return True

    def run(self, lockstep=False):
        """Persistent grid: `resident` CTAs; a CTA takes the next ticket when its tile is finished.
        lockstep=False: every scheduler step advances one randomly chosen CTA by one phase (compute ->
        publish -> trace-ahead -> look-back): arbitrary interleavings.  lockstep=True: sweeps over all CTAs,
        each advancing with probability 0.9 -- the resident wave moves through its phases together, as it
        does on the device (all CTAs start at once and do the same work)."""
        n = len(self.totals)
        nxt = 0
        ctas = []
        for _ in range(min(self.resident, n)):
            ctas.append([nxt, 0])
            nxt += 1
        steps = 0
        while ctas:
            steps += 1
            assert steps < 400 * n + 10000, "no progress: the protocol deadlocked"
            if lockstep:
                batch = [c for c in ctas if self.rng.random() < 0.9]
                self.rng.shuffle(batch)
            else:
                batch = [self.rng.choice(ctas)]
            for c in batch:
                t, phase = c
                if phase < 2:                      # material / scan work of varying length
                    c[1] += 1 if (lockstep or self.rng.random() < 0.5) else 0
                elif phase == 2:
                    self.publish(t)
                    c[1] = 3
                elif phase < 5:                    # trace-ahead work
                    c[1] += 1 if (lockstep or self.rng.random() < 0.5) else 0
                else:
                    if self.lookback(t):
                        if nxt < n:
                            c[0], c[1] = nxt, 0
                            nxt += 1
                        else:
                            ctas.remove(c)
        return self


@pytest.mark.parametrize("grouped", [False, True], ids=["flat", "grouped"])
@pytest.mark.parametrize("n_tiles,resident,seed", [(1, 4, 0), (31, 8, 1), (32, 8, 2), (33, 64, 3), (700, 96, 4),
                                                    (1500, 592, 5), (2049, 37, 6)])
def test_lookback_yields_exact_exclusive_prefix(grouped, n_tiles, resident, seed):
    rng = random.Random(100 + seed)
    totals = [rng.choice((0, 1, 77, 128, 200, 256)) for _ in range(n_tiles)]
    sim = Sim(totals, grouped, resident, seed).run(lockstep=bool(seed % 2))
    acc = 0
    for t in range(n_tiles):
        assert sim.prefix[t] == acc, "tile %d" % t
        acc += totals[t]
