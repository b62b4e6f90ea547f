#!/usr/bin/env python
"""Discrete-event model of the mbarrier protocol of the warp-specialised pass A (paw::pass_a_ws_kernel in
sobfu_b200/csrc/solver_tiled.cu): one feeder lane (inside sampler warp 0), NSAMP sampler warps, NWS stencil warps, the psi ring
(full / empty barriers, TMA completion modelled as an asynchronous event) and the w ring (wfull / wempty).  Random interleavings
are explored; the model checks that nothing deadlocks, that every read sees the plane it expects (no phase lapping) and that no
buffer is overwritten while a reader still needs it.  It mirrors the index arithmetic of the kernel line by line, so it is a check
of the protocol, not of the CUDA details.  Usage: python tools/sim_pass_a_ws.py [runs] [positions]"""
import random
import sys

NSTAGE, AHEAD, NWB, NWS, NSAMP = 6, 2, 5, 4, 8


class Bar:
    """mbarrier: `count` arrivals (+ optional transaction) complete a phase; wait(parity) passes once the phase with that parity
    has completed, i.e. while the number of completed phases is odd for parity 0, even (and > 0 ... by construction) for parity 1"""

    def __init__(self, count):
        self.count, self.pending, self.done, self.tx = count, count, 0, False

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier expects in one phase"
        self._maybe()

    def expect_tx(self):      # arrive.expect_tx: one arrival + an outstanding transaction
        self.tx = True
        self.arrive()

    def complete_tx(self):
        self.tx = False
        self._maybe()

    def _maybe(self):
        if self.pending == 0 and not self.tx:
            self.done += 1
            self.pending = self.count

    def passed(self, parity):
        return (self.done & 1) != parity


def run(seed, Q):
    rng = random.Random(seed)
    full = [Bar(1) for _ in range(NSTAGE)]
    empty = [Bar(NWS + NSAMP) for _ in range(NSTAGE)]
    wfull = [Bar(NSAMP) for _ in range(NWB)]
    wempty = [Bar(NWS) for _ in range(NWB)]
    psi = [None] * NSTAGE                     # position whose plane the stage holds
    w = [[None] * NSAMP for _ in range(NWB)]  # per sampler warp's share
    samp_done = [0] * NSAMP                   # positions finished by each sampler warp
    sten_done = [0] * NWS                     # steps finished by each stencil warp
    inflight = []                             # (slot, position) TMA loads issued, not yet landed

    def sampler(k):
        qi = 0
        for q in range(Q):
            if k == 0:                        # the feeder lane runs first in warp 0's step
                while qi <= q + AHEAD and qi < Q:
                    slot, n = qi % NSTAGE, qi // NSTAGE
                    if n > 0:
                        while not empty[slot].passed((n - 1) & 1):
                            yield
                        old = qi - NSTAGE     # nobody may still need the plane that is overwritten
                        assert all(d > old for d in samp_done) and all(d > old + 2 or d >= Q for d in sten_done), ("psi overwrite", qi, samp_done, sten_done)
                    full[slot].expect_tx()
                    inflight.append((slot, qi))
                    qi += 1
            slot, wb = q % NSTAGE, q % NWB
            while not full[slot].passed((q // NSTAGE) & 1):
                yield
            if q >= NWB:
                while not wempty[wb].passed(((q // NWB) - 1) & 1):
                    yield
                assert all(d > q - NWB + 2 or d >= Q for d in sten_done), ("w overwrite", q, sten_done)
            assert psi[slot] == q, ("sampler reads wrong plane", k, q, psi[slot])
            yield                             # sampling takes time
            w[wb][k] = q
            wfull[wb].arrive()
            empty[slot].arrive()
            samp_done[k] = q + 1
            yield

    def stencil(k):
        q = 0
        for step in range(Q):
            slot, wb = q % NSTAGE, q % NWB
            while not full[slot].passed((q // NSTAGE) & 1):
                yield
            while not wfull[wb].passed((q // NWB) & 1):
                yield
            q += 1
            if q >= 3:                        # centre_on: planes q-3, q-2, q-1 of the stream
                for back in (1, 2, 3):
                    pos = q - back
                    assert psi[pos % NSTAGE] == pos, ("stencil reads wrong psi", k, pos, psi[pos % NSTAGE])
                    assert all(t == pos for t in w[pos % NWB]), ("stencil reads wrong w", k, pos, w[pos % NWB])
            yield                             # the stencil arithmetic
            sten_done[k] = q
            if q >= 3:
                empty[(q - 3) % NSTAGE].arrive()
                wempty[(q - 3) % NWB].arrive()
            yield

    agents = [sampler(k) for k in range(NSAMP)] + [stencil(k) for k in range(NWS)]
    alive = list(range(len(agents)))
    idle = 0
    while alive:
        # TMA loads land at random times, in any order
        if inflight and rng.random() < 0.5:
            slot, pos = inflight.pop(rng.randrange(len(inflight)))
            psi[slot] = pos
            full[slot].complete_tx()
        a = rng.choice(alive)
        before = (tuple(b.done for b in full + empty + wfull + wempty), tuple(samp_done), tuple(sten_done), len(inflight))
        try:
            next(agents[a])
        except StopIteration:
            alive.remove(a)
        after = (tuple(b.done for b in full + empty + wfull + wempty), tuple(samp_done), tuple(sten_done), len(inflight))
        idle = idle + 1 if before == after else 0
        assert idle < 20000, ("deadlock", samp_done, sten_done, [b.done for b in full], [b.done for b in empty])
    assert all(d == Q for d in samp_done) and all(d == Q for d in sten_done)


if __name__ == "__main__":
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    Q = int(sys.argv[2]) if len(sys.argv) > 2 else 45
    for seed in range(runs):
        run(seed, Q)
    print("ok: %d random interleavings of %d stream positions, no deadlock, no stale read, no early overwrite" % (runs, Q))
