#!/usr/bin/env python
"""Host-side model of the synchronisation protocol of ``csrc/dftf4.cu`` (the dual-tile STFT GEMM).

The kernel could not be run when it was written (no GPU budget left), so its barrier protocol is checked here instead: the
producer warps of both CTAs, the MMA issuer, the 16 epilogue warps, the TMA engine and the tensor pipe are coroutines that
follow the kernel's loops line by line over modelled mbarriers (arrival counts, transaction bytes, phase parity), shared-memory
ring slots and TMEM regions, under a randomised scheduler.  The model fails on

* a deadlock (no agent can make progress before everybody is done),
* a ring slot refilled before the MMAs reading it have retired, or read before its load has landed,
* a TMEM region overwritten before every epilogue warp has drained it, or drained before its accumulator is complete,
* an epilogue warp combining Re / Im of different tiles, parts or frame pairs,
* an mbarrier running more than one phase ahead of a waiter (parity aliasing).

    python tools/dual_protocol_sim.py [pairs_per_cluster] [seeds]

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
    def __init__(self, pairs, seed):
        self.rng = random.Random(seed)
        self.pairs = pairs
        B = Barrier
        # leader-resident barriers (index 0) and per-CTA barriers (index by cta)
        self.a_full = [B(f"a_full{s}", 2) for s in range(K_SA)]
        self.b_full = [B(f"b_full{s}", 2) for s in range(K_SB)]
        self.a_empty = [[B(f"a_empty{c}.{s}", 1) for s in range(K_SA)] for c in range(2)]
        self.b_empty = [[B(f"b_empty{c}.{s}", 1) for s in range(K_SB)] for c in range(2)]
        self.r_full = [[B(f"r_full{c}.{r}", 1) for r in range(REGIONS)] for c in range(2)]
        self.r_empty = [B(f"r_empty{r}", 2 * EPI_WARPS) for r in range(REGIONS)]
        # data: what each ring slot holds (per CTA) and what each TMEM region holds (shared view of the pair)
        self.slot_a = [[None] * K_SA for _ in range(2)]
        self.slot_b = [[None] * K_SB for _ in range(2)]
        self.slot_a_readers = [[0] * K_SA for _ in range(2)]      # MMAs issued but not yet retired that read the slot
        self.slot_b_readers = [[0] * K_SB for _ in range(2)]
        self.region = [dict(tag=None, kblocks=0, complete=False, drained=2 * EPI_WARPS) for _ in range(REGIONS)]
        self.tma_q, self.mma_q = [], []
        self.outputs = []

    # ------------------------------------------------------------------ asynchronous engines
    def tma_engine(self):
        while True:
            if self.tma_q and self.rng.random() < 0.7:
                i = self.rng.randrange(min(3, len(self.tma_q)))          # loads may land out of order
                cta, ring, slot, content, bar, nbytes = self.tma_q.pop(i)
                slots = self.slot_a if ring == "a" else self.slot_b
                readers = self.slot_a_readers if ring == "a" else self.slot_b_readers
                assert readers[cta][slot] == 0, f"ring {ring} slot {slot} of CTA {cta} refilled while MMAs still read it"
                slots[cta][slot] = content
                bar.complete_tx(nbytes)
            yield

    def tensor_pipe(self):
        while True:
            if self.mma_q and self.rng.random() < 0.6:
                op = self.mma_q.pop(0)                                   # in order
                if op[0] == "mma":
                    _, sa, sb, r, tag, kb, first, a_want, b_want = op
                    for cta in range(2):
                        assert self.slot_a[cta][sa] == a_want, f"A slot {sa}: holds {self.slot_a[cta][sa]}, MMA wants {a_want}"
                        assert self.slot_b[cta][sb] == b_want, f"B slot {sb}: holds {self.slot_b[cta][sb]}, MMA wants {b_want}"
                        self.slot_a_readers[cta][sa] -= 1
                        self.slot_b_readers[cta][sb] -= 1
                    reg = self.region[r]
                    if first:
                        assert reg["drained"] == 2 * EPI_WARPS, f"region {r} overwritten by {tag} before {reg['tag']} was drained"
                        reg.update(tag=tag, kblocks=0, complete=False, drained=0)
                    assert reg["tag"] == tag and reg["kblocks"] == kb, f"region {r}: accumulating {tag} kb {kb} onto {reg}"
                    reg["kblocks"] += 1
                else:
                    _, bars, done_region = op
                    if done_region is not None:
                        self.region[done_region]["complete"] = True
                    for b in bars:
                        b.arrive()
            yield

    # ------------------------------------------------------------------ kernel agents (follow dftf4.cu)
    def producer(self, cta):
        sa = sb = 0
        pa = pb = 0
        for pair in range(self.pairs):
            for g, G in enumerate(GROUPS):
                kbp = G["kbp"]
                for kb in range(2 * kbp):
                    part = 0 if kb < kbp else 1
                    while not self.a_empty[cta][sa].passed(pa ^ 1, ("prod", cta), "empty"):
                        yield
                    if cta == 0:
                        self.a_full[sa].arrive(expect_tx=2)
                    else:
                        self.a_full[sa].arrive()
                    self.tma_q.append((cta, "a", sa, ("A", pair, g, kb), self.a_full[sa], 1))
                    sa += 1
                    if sa == K_SA:
                        sa, pa = 0, pa ^ 1
                    for t in range(G["tiles"]):
                        while not self.b_empty[cta][sb].passed(pb ^ 1, ("prod", cta), "empty"):
                            yield
                        if cta == 0:
                            self.b_full[sb].arrive(expect_tx=2)
                        else:
                            self.b_full[sb].arrive()
                        self.tma_q.append((cta, "b", sb, ("B", G["item0"] + t, part, kb - part * kbp), self.b_full[sb], 1))
                        sb += 1
                        if sb == K_SB:
                            sb, pb = 0, pb ^ 1
                    yield

    def issuer(self):
        sa = sb = 0
        pa = pb = 0
        used = 0
        for pair in range(self.pairs):
            for g, G in enumerate(GROUPS):
                kbp = G["kbp"]
                for part in range(2):
                    for kb in range(kbp):
                        while not self.a_full[sa].passed(pa, "issuer", "full"):
                            yield
                        for t in range(G["tiles"]):
                            r = region_of(g, t, part)
                            if kb == 0:
                                while not self.r_empty[r].passed(((used >> r) & 1) ^ 1, "issuer", "empty"):
                                    yield
                            while not self.b_full[sb].passed(pb, "issuer", "full"):
                                yield
                            tag = (pair, G["item0"] + t, part)
                            for cta in range(2):
                                self.slot_a_readers[cta][sa] += 1
                                self.slot_b_readers[cta][sb] += 1
                            self.mma_q.append(("mma", sa, sb, r, tag, kb, kb == 0, ("A", pair, g, part * kbp + kb),
                                               ("B", G["item0"] + t, part, kb)))
                            self.mma_q.append(("commit", [self.b_empty[0][sb], self.b_empty[1][sb]], None))
                            if kb == kbp - 1:
                                self.mma_q.append(("commit", [self.r_full[0][r], self.r_full[1][r]], r))
                                used ^= 1 << r
                            sb += 1
                            if sb == K_SB:
                                sb, pb = 0, pb ^ 1
                            yield
                        self.mma_q.append(("commit", [self.a_empty[0][sa], self.a_empty[1][sa]], None))
                        sa += 1
                        if sa == K_SA:
                            sa, pa = 0, pa ^ 1
                        yield

    def epilogue(self, cta, w):
        used = 0
        for pair in range(self.pairs):
            for g, G in enumerate(GROUPS):
                for t in range(G["tiles"]):
                    it = G["item0"] + t
                    got = []
                    for part in range(2):
                        r = epilogue_region_of(g, t, part)
                        while not self.r_full[cta][r].passed((used >> r) & 1, ("epi", cta, w), "full"):
                            yield
                        reg = self.region[r]
                        assert reg["complete"] and reg["tag"] == (pair, it, part) and reg["kblocks"] == G["kbp"], \
                            f"epilogue ({cta},{w}) expected {(pair, it, part)} in region {r}, found {reg}"
                        got.append(reg["tag"])
                        yield                                             # the tcgen05.ld takes a while
                        assert reg["tag"] == (pair, it, part), f"region {r} changed under a read of {(pair, it, part)}"
                        reg["drained"] += 1
                        self.r_empty[r].arrive()
                        used ^= 1 << r
                        yield
                    self.outputs.append((cta, w, pair, it, tuple(got)))

    # ------------------------------------------------------------------ scheduler
    def run(self):
        agents = {"tma": self.tma_engine(), "pipe": self.tensor_pipe(), "issuer": self.issuer()}
        for c in range(2):
            agents[f"prod{c}"] = self.producer(c)
            for w in range(EPI_WARPS):
                agents[f"epi{c}.{w}"] = self.epilogue(c, w)
        live = {k for k in agents if k not in ("tma", "pipe")}
        idle_rounds = 0
        rounds = 0
        speed = {}
        while live:
            if rounds % 150 == 0:                                         # adversarial pacing: every agent is at times
                speed = {n: self.rng.choice((1.0, 1.0, 0.3, 0.03)) for n in agents}     # much slower than the others
            rounds += 1
            names = list(agents)
            self.rng.shuffle(names)
            before = (len(self.tma_q), len(self.mma_q), len(self.outputs), tuple(b.phase for b in self.all_barriers()))
            for n in names:
                if (n in live or n in ("tma", "pipe")) and self.rng.random() < speed[n]:
                    try:
                        next(agents[n])
                    except StopIteration:
                        live.discard(n)
            after = (len(self.tma_q), len(self.mma_q), len(self.outputs), tuple(b.phase for b in self.all_barriers()))
            idle_rounds = idle_rounds + 1 if before == after else 0
            assert idle_rounds < 20000, f"deadlock: waiting agents {sorted(live)}"
        while self.mma_q or self.tma_q:                                   # drain the engines
            next(agents["tma"])
            next(agents["pipe"])
        want = 2 * EPI_WARPS * self.pairs * 4
        assert len(self.outputs) == want, f"{len(self.outputs)} tile epilogues, expected {want}"
        return True

    def all_barriers(self):
        out = self.a_full + self.b_full + self.r_empty
        for c in range(2):
            out += self.a_empty[c] + self.b_empty[c] + self.r_full[c]
        return out


def main(argv):
    pairs = int(argv[1]) if len(argv) > 1 else 4
    seeds = int(argv[2]) if len(argv) > 2 else 20
    for seed in range(seeds):
        Sim(pairs, seed).run()
    print(f"dual-tile protocol: {seeds} randomised schedules x {pairs} frame pairs per cluster -- no deadlock, no hazard")


if __name__ == "__main__":
    main(sys.argv)
