#!/usr/bin/env python
"""Host-side model of the synchronisation protocol of ``csrc/dftf4.cu`` (the dual-tile STFT GEMM).

The kernel could not be run when it was written (no GPU budget left), so its barrier protocol is checked here instead: the
producer warps of both CTAs, the MMA issuer, the 16 epilogue warps of every CTA pair, the TMA engine and the tensor pipes are coroutines that
follow the kernel's loops line by line over modelled mbarriers (arrival counts, transaction bytes, phase parity), shared-memory
ring slots and TMEM regions, under a randomised scheduler.  The model fails on

* a deadlock (no agent can make progress before everybody is done),
* a ring slot refilled before the MMAs reading it have retired, or read before its load has landed,
* a TMEM region overwritten before every epilogue warp has drained it, or drained before its accumulator is complete,
* an epilogue warp combining Re / Im of different tiles, parts or frame pairs,
* an mbarrier running more than one phase ahead of a waiter (parity aliasing).

    python tools/dual_protocol_sim.py [rounds] [seeds]        (both variants: 1 and 2 CTA pairs per cluster)

What it does NOT cover: anything about the instructions themselves (descriptors, TMEM addressing, swizzles, fences).
"""
from __future__ import annotations

import random
import sys

K_SA, K_SB, REGIONS, EPI_WARPS = 4, 4, 3, 8
GROUPS = [dict(kbp=8, item0=0, tiles=2), dict(kbp=4, item0=2, tiles=1), dict(kbp=4, item0=3, tiles=1)]


def region_of(g, t, part):                      # dftf4.cu::region_of
    if g == 0:
        return (0 if part == 0 else 2) if t == 0 else (1 if part == 0 else 0)
    if g == 1:
        return 2 if part == 0 else 1
    return 0 if part == 0 else 2


epilogue_region_of = region_of                  # the epilogue calls the same function (negative controls replace this one)


class Barrier:
    def __init__(self, name, count):
        self.name, self.count = name, count
        self.phase, self.pending, self.tx = 0, count, 0
        self.waiters_seen = {}

    def _maybe_complete(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self, expect_tx=0):
        assert self.pending > 0, f"{self.name}: more arrivals than the barrier expects"
        self.tx += expect_tx
        self.pending -= 1
        self._maybe_complete()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._maybe_complete()

    def passed(self, parity, who, kind):
        """mbarrier.try_wait.parity: true once the phase with this parity has completed.  ``kind`` says what the waiter means:
        its n-th wait (n = 0, 1, ...) on a "full" barrier needs completion n + 1, on an "empty" barrier completion n (the
        first pass through a ring finds every slot free).  If the barrier is further ahead than that when the wait succeeds,
        the parity has been reused and the waiter has sailed through a phase it never saw."""
        if (self.phase & 1) == parity:
            return False
        n = self.waiters_seen.get(who, 0)
        want = n + 1 if kind == "full" else n
        assert self.phase == want, f"{self.name}: waiter {who} passes wait #{n} ({kind}) at phase {self.phase}, expected {want}"
        self.waiters_seen[who] = n + 1
        return True


class Sim:
    """``npairs`` CTA pairs per cluster: 1 = AVLD_DFT_DUAL=1, 2 = the B-multicast variant (AVLD_DFT_DUAL=2)."""

    def __init__(self, rounds, seed, npairs=1):
        self.rng = random.Random(seed)
        self.rounds, self.P = rounds, npairs
        B, P = Barrier, npairs
        rng2, rngP = range(2), range(P)
        # per pair: leader-resident barriers; per pair and CTA: the barriers multicast commits arrive on
        self.a_full = [[B(f"a_full{q}.{s}", 2) for s in range(K_SA)] for q in rngP]
        self.b_full = [[B(f"b_full{q}.{s}", 2) for s in range(K_SB)] for q in rngP]
        self.a_empty = [[[B(f"a_empty{q}.{c}.{s}", 1) for s in range(K_SA)] for c in rng2] for q in rngP]
        self.b_empty = [[[B(f"b_empty{q}.{c}.{s}", P) for s in range(K_SB)] for c in rng2] for q in rngP]   # one commit per issuer
        self.r_full = [[[B(f"r_full{q}.{c}.{r}", 1) for r in range(REGIONS)] for c in rng2] for q in rngP]
        self.r_empty = [[B(f"r_empty{q}.{r}", 2 * EPI_WARPS) for r in range(REGIONS)] for q in rngP]
        # data: ring slots per pair and CTA (a B slot is filled in P shares, one per loading pair), TMEM regions per pair
        self.slot_a = [[[None] * K_SA for _ in rng2] for _ in rngP]
        self.slot_b = [[[[None] * P for _ in range(K_SB)] for _ in rng2] for _ in rngP]
        self.slot_a_readers = [[[0] * K_SA for _ in rng2] for _ in rngP]      # MMAs issued but not retired that read the slot
        self.slot_b_readers = [[[0] * K_SB for _ in rng2] for _ in rngP]
        self.region = [[dict(tag=None, kblocks=0, complete=False, drained=2 * EPI_WARPS) for _ in range(REGIONS)] for _ in rngP]
        self.tma_q = []
        self.mma_q = [[] for _ in rngP]
        self.outputs = []

    # ------------------------------------------------------------------ asynchronous engines
    def tma_engine(self):
        while True:
            if self.tma_q and self.rng.random() < 0.7:
                i = self.rng.randrange(min(4, len(self.tma_q)))          # loads may land out of order
                q, cta, ring, slot, share, content, bar = self.tma_q.pop(i)
                if ring == "a":
                    assert self.slot_a_readers[q][cta][slot] == 0, f"A slot {slot} of CTA {q}.{cta} refilled while MMAs still read it"
                    self.slot_a[q][cta][slot] = content
                else:
                    assert self.slot_b_readers[q][cta][slot] == 0, f"B slot {slot} of CTA {q}.{cta} refilled while MMAs still read it"
                    self.slot_b[q][cta][slot][share] = content
                bar.complete_tx(1)
            yield

    def tensor_pipe(self, q):
        while True:
            if self.mma_q[q] and self.rng.random() < 0.6:
                op = self.mma_q[q].pop(0)                                # in order
                if op[0] == "mma":
                    _, sa, sb, r, tag, kb, first, a_want, b_want = op
                    for cta in range(2):
                        assert self.slot_a[q][cta][sa] == a_want, f"A slot {sa}: holds {self.slot_a[q][cta][sa]}, MMA wants {a_want}"
                        assert self.slot_b[q][cta][sb] == [b_want] * self.P, \
                            f"B slot {sb} of CTA {q}.{cta}: holds {self.slot_b[q][cta][sb]}, MMA wants {b_want}"
                        self.slot_a_readers[q][cta][sa] -= 1
                        self.slot_b_readers[q][cta][sb] -= 1
                    reg = self.region[q][r]
                    if first:
                        assert reg["drained"] == 2 * EPI_WARPS, f"region {r} overwritten by {tag} before {reg['tag']} was drained"
                        reg.update(tag=tag, kblocks=0, complete=False, drained=0)
                    assert reg["tag"] == tag and reg["kblocks"] == kb, f"region {r}: accumulating {tag} kb {kb} onto {reg}"
                    reg["kblocks"] += 1
                else:
                    _, bars, done_region = op
                    if done_region is not None:
                        self.region[q][done_region]["complete"] = True
                    for b in bars:
                        b.arrive()
            yield

    # ------------------------------------------------------------------ kernel agents (follow dftf4.cu)
    def producer(self, q, cta):
        sa = sb = 0
        pa = pb = 0
        P = self.P
        for rnd in range(self.rounds):
            pair = rnd * P + q
            for g, G in enumerate(GROUPS):
                kbp = G["kbp"]
                for kb in range(2 * kbp):
                    part = 0 if kb < kbp else 1
                    while not self.a_empty[q][cta][sa].passed(pa ^ 1, ("prod", q, cta), "empty"):
                        yield
                    if cta == 0:
                        self.a_full[q][sa].arrive(expect_tx=2)
                    else:
                        self.a_full[q][sa].arrive()
                    self.tma_q.append((q, cta, "a", sa, 0, ("A", pair, g, kb), self.a_full[q][sa]))
                    sa += 1
                    if sa == K_SA:
                        sa, pa = 0, pa ^ 1
                    for t in range(G["tiles"]):
                        while not self.b_empty[q][cta][sb].passed(pb ^ 1, ("prod", q, cta), "empty"):
                            yield
                        if cta == 0:
                            self.b_full[q][sb].arrive(expect_tx=2 * P)
                        else:
                            self.b_full[q][sb].arrive()
                        content = ("B", G["item0"] + t, part, kb - part * kbp)
                        for q2 in range(P):                              # multicast: share q lands in the same-parity CTA of every pair
                            self.tma_q.append((q2, cta, "b", sb, q, content, self.b_full[q2][sb]))
                        sb += 1
                        if sb == K_SB:
                            sb, pb = 0, pb ^ 1
                    yield

    def issuer(self, q):
        sa = sb = 0
        pa = pb = 0
        used = 0
        P = self.P
        for rnd in range(self.rounds):
            pair = rnd * P + q
            for g, G in enumerate(GROUPS):
                kbp = G["kbp"]
                for part in range(2):
                    for kb in range(kbp):
                        while not self.a_full[q][sa].passed(pa, ("issuer", q), "full"):
                            yield
                        for t in range(G["tiles"]):
                            r = region_of(g, t, part)
                            if kb == 0:
                                while not self.r_empty[q][r].passed(((used >> r) & 1) ^ 1, ("issuer", q), "empty"):
                                    yield
                            while not self.b_full[q][sb].passed(pb, ("issuer", q), "full"):
                                yield
                            tag = (pair, G["item0"] + t, part)
                            for cta in range(2):
                                self.slot_a_readers[q][cta][sa] += 1
                                self.slot_b_readers[q][cta][sb] += 1
                            self.mma_q[q].append(("mma", sa, sb, r, tag, kb, kb == 0, ("A", pair, g, part * kbp + kb),
                                                  ("B", G["item0"] + t, part, kb)))
                            self.mma_q[q].append(("commit", [self.b_empty[q2][c][sb] for q2 in range(P) for c in range(2)], None))
                            if kb == kbp - 1:
                                self.mma_q[q].append(("commit", [self.r_full[q][0][r], self.r_full[q][1][r]], r))
                                used ^= 1 << r
                            sb += 1
                            if sb == K_SB:
                                sb, pb = 0, pb ^ 1
                            yield
                        self.mma_q[q].append(("commit", [self.a_empty[q][0][sa], self.a_empty[q][1][sa]], None))
                        sa += 1
                        if sa == K_SA:
                            sa, pa = 0, pa ^ 1
                        yield

    def epilogue(self, q, cta, w):
        used = 0
        for rnd in range(self.rounds):
            pair = rnd * self.P + q
            for g, G in enumerate(GROUPS):
                for t in range(G["tiles"]):
                    it = G["item0"] + t
                    got = []
                    for part in range(2):
                        r = epilogue_region_of(g, t, part)
                        while not self.r_full[q][cta][r].passed((used >> r) & 1, ("epi", q, cta, w), "full"):
                            yield
                        reg = self.region[q][r]
                        assert reg["complete"] and reg["tag"] == (pair, it, part) and reg["kblocks"] == G["kbp"], \
                            f"epilogue ({q},{cta},{w}) expected {(pair, it, part)} in region {r}, found {reg}"
                        got.append(reg["tag"])
                        yield                                             # the tcgen05.ld takes a while
                        assert reg["tag"] == (pair, it, part), f"region {r} changed under a read of {(pair, it, part)}"
                        reg["drained"] += 1
                        self.r_empty[q][r].arrive()
                        used ^= 1 << r
                        yield
                    self.outputs.append((q, cta, w, pair, it, tuple(got)))

    # ------------------------------------------------------------------ scheduler
    def run(self):
        agents = {"tma": self.tma_engine()}
        for q in range(self.P):
            agents[f"pipe{q}"] = self.tensor_pipe(q)
            agents[f"issuer{q}"] = self.issuer(q)
            for c in range(2):
                agents[f"prod{q}.{c}"] = self.producer(q, c)
                for w in range(EPI_WARPS):
                    agents[f"epi{q}.{c}.{w}"] = self.epilogue(q, c, w)
        engines = {n for n in agents if n == "tma" or n.startswith("pipe")}
        live = set(agents) - engines
        idle_rounds = 0
        rounds = 0
        speed = {}
        while live:
            if rounds % 150 == 0:                                         # adversarial pacing: every agent is at times
                speed = {n: self.rng.choice((1.0, 1.0, 0.3, 0.03)) for n in agents}     # much slower than the others
            rounds += 1
            names = list(agents)
            self.rng.shuffle(names)
            before = (len(self.tma_q), sum(map(len, self.mma_q)), len(self.outputs), tuple(b.phase for b in self.all_barriers()))
            for n in names:
                if (n in live or n in engines) and self.rng.random() < speed[n]:
                    try:
                        next(agents[n])
                    except StopIteration:
                        live.discard(n)
            after = (len(self.tma_q), sum(map(len, self.mma_q)), len(self.outputs), tuple(b.phase for b in self.all_barriers()))
            idle_rounds = idle_rounds + 1 if before == after else 0
            assert idle_rounds < 20000, f"deadlock: waiting agents {sorted(live)}"
        while self.tma_q or any(self.mma_q):                              # drain the engines
            for n in engines:
                next(agents[n])
        want = self.P * 2 * EPI_WARPS * self.rounds * 4
        assert len(self.outputs) == want, f"{len(self.outputs)} tile epilogues, expected {want}"
        return True

    def all_barriers(self):
        out = []
        for q in range(self.P):
            out += self.a_full[q] + self.b_full[q] + self.r_empty[q]
            for c in range(2):
                out += self.a_empty[q][c] + self.b_empty[q][c] + self.r_full[q][c]
        return out


def main(argv):
    rounds = int(argv[1]) if len(argv) > 1 else 4
    seeds = int(argv[2]) if len(argv) > 2 else 20
    for npairs in (1, 2):
        for seed in range(seeds):
            Sim(rounds, seed, npairs).run()
        print(f"dual-tile protocol, {npairs} CTA pair(s) per cluster: {seeds} randomised schedules x {rounds} rounds -- "
              f"no deadlock, no hazard")


if __name__ == "__main__":
    main(sys.argv)
