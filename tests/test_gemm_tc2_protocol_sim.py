"""Discrete-event simulation of the barrier protocol of the EXPERIMENTAL CTA-pair GEMM (csrc/gemm_tc2.cu): two CTAs, one MMA
issuer (leader), both producers crediting the LEADER's full barrier (expect_tx armed by the leader only, possibly after the
peer's bytes have already landed), multicast commits releasing the stage / publishing the accumulator in both CTAs, remote
arrivals of the peer's epilogue warps on the leader's tempty barrier.  Same approach as test_attn_v2_protocol_sim.py."""
import heapq
import random

import pytest

STAGES = 6


class TxBar:
    """mbarrier with a transaction count: the phase completes when all expected arrivals happened AND tx == 0"""

    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.pending = self.count
            self.phase += 1

    def arrive(self):
        assert self.pending > 0, "more arrivals than expected in one phase"
        self.pending -= 1
        self._check()

    def arrive_expect_tx(self, nbytes):
        assert self.pending > 0
        self.tx += nbytes
        self.pending -= 1
        self._check()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._check()


class Buf:
    def __init__(self, name):
        self.name, self.readers, self.writer, self.content = name, 0, None, None

    def write_begin(self, tag):
        assert self.readers == 0 and self.writer is None, f"{self.name}: write {tag} while busy ({self.content})"
        self.writer = tag

    def write_end(self, tag):
        assert self.writer == tag
        self.writer, self.content = None, tag

    def read_begin(self, tag):
        assert self.writer is None and self.content == tag, f"{self.name}: read {tag} sees {self.content} / writer {self.writer}"
        self.readers += 1

    def read_end(self):
        self.readers -= 1


class PairSim:
    def __init__(self, ntiles, num_kb, n_epi, seed, leader_arms_late=False):
        self.rng = random.Random(seed)
        self.now, self.events, self.seq = 0.0, [], 0
        self.ntiles, self.num_kb, self.n_epi, self.late = ntiles, num_kb, n_epi, leader_arms_late
        self.full = [TxBar(1) for _ in range(STAGES)]                       # leader only
        self.empty = [[TxBar(1) for _ in range(STAGES)] for _ in range(2)]  # per CTA, multicast commit
        self.tfull = [[TxBar(1) for _ in range(2)] for _ in range(2)]       # per CTA, multicast commit
        self.tempty = [TxBar(2 * n_epi) for _ in range(2)]                  # leader only, arrivals from both CTAs
        self.stage = [[Buf(f"cta{c}.stage{s}") for s in range(STAGES)] for c in range(2)]
        self.acc = [[Buf(f"cta{c}.acc{a}") for a in range(2)] for c in range(2)]
        self.mma_queue, self.mma_busy = [], False
        self.stored = [[0] * ntiles for _ in range(2)]

    def at(self, dt, fn):
        self.seq += 1
        heapq.heappush(self.events, (self.now + dt, self.seq, fn))

    def tma(self, buf, tag, bar, nbytes):
        buf.write_begin(tag)

        def done():
            buf.write_end(tag)
            bar.complete_tx(nbytes)
        self.at(self.rng.uniform(0.2, 4.0), done)

    def mma(self, reads, writes, rtag, wtag, commits):
        self.mma_queue.append((reads, writes, rtag, wtag, commits))
        self._pump()

    def _pump(self):
        if self.mma_busy or not self.mma_queue:
            return
        reads, writes, rtag, wtag, commits = self.mma_queue.pop(0)
        self.mma_busy = True
        for b in reads:
            b.read_begin(rtag)
        for w in writes:
            w.write_begin(wtag)

        def done():
            for b in reads:
                b.read_end()
            for w in writes:
                w.write_end(wtag)
            for bar in commits:
                bar.arrive()
            self.mma_busy = False
            self._pump()
        self.at(self.rng.uniform(0.05, 0.8), done)

    # ---- roles
    def producer(self, cta):
        it = 0
        for tile in range(self.ntiles):
            for kb in range(self.num_kb):
                s, ph = it % STAGES, (it // STAGES) & 1
                yield ("wait", self.empty[cta][s], ph ^ 1)
                if cta == 0:
                    if self.late:
                        yield ("delay", self.rng.uniform(0.0, 6.0))       # the peer's bytes may land before the barrier is armed
                    self.full[s].arrive_expect_tx(2 * 100)
                self.tma(self.stage[cta][s], (tile, kb), self.full[s], 100)
                yield ("delay", self.rng.uniform(0.0, 0.2))
                it += 1

    def mma_thread(self):
        it = 0
        for tl in range(self.ntiles):
            acc, aph = tl & 1, (tl >> 1) & 1
            yield ("wait", self.tempty[acc], aph ^ 1)
            for kb in range(self.num_kb):
                s, ph = it % STAGES, (it // STAGES) & 1
                yield ("wait", self.full[s], ph)
                self.mma([self.stage[0][s], self.stage[1][s]], [self.acc[0][acc], self.acc[1][acc]], (tl, kb), tl,
                         [self.empty[0][s], self.empty[1][s]] + ([self.tfull[0][acc], self.tfull[1][acc]] if kb == self.num_kb - 1 else []))
                it += 1
                yield ("delay", self.rng.uniform(0.0, 0.1))

    def epilogue(self, cta, w):
        for tl in range(self.ntiles):
            acc, aph = tl & 1, (tl >> 1) & 1
            yield ("wait", self.tfull[cta][acc], aph)
            self.acc[cta][acc].read_begin(tl)
            yield ("delay", self.rng.uniform(0.3, 5.0))
            self.acc[cta][acc].read_end()
            if w == 0:
                self.stored[cta][tl] += 1
            self.tempty[acc].arrive()            # remote arrive on the leader's barrier when cta == 1

    def run(self):
        agents = [self.producer(0), self.producer(1), self.mma_thread()] + [self.epilogue(c, w) for c in range(2) for w in range(self.n_epi)]
        blocked, alive = {}, set(range(len(agents)))

        def step(i):
            try:
                req = next(agents[i])
            except StopIteration:
                alive.discard(i)
                return
            if req[0] == "delay":
                self.at(req[1], lambda i=i: step(i))
            else:
                blocked[i] = (req[1], req[2])

        for i in list(alive):
            step(i)
        guard = 0
        while alive:
            guard += 1
            assert guard < 400000, "simulation does not terminate"
            progressed = False
            for i, (bar, parity) in list(blocked.items()):
                if (bar.phase & 1) != parity:
                    del blocked[i]
                    step(i)
                    progressed = True
            if progressed:
                continue
            if not self.events:
                raise AssertionError(f"deadlock at t={self.now:.2f}: blocked agents {sorted(blocked)}")
            self.now, _, fn = heapq.heappop(self.events)
            fn()
        while self.events:
            self.now, _, fn = heapq.heappop(self.events)
            fn()
        assert self.stored == [[1] * self.ntiles, [1] * self.ntiles]


@pytest.mark.parametrize("ntiles,num_kb", [(1, 3), (2, 20), (5, 7), (9, 60)])
@pytest.mark.parametrize("late", [False, True])
def test_pair_protocol_terminates_without_hazards(ntiles, num_kb, late):
    for seed in range(25):
        PairSim(ntiles, num_kb, n_epi=2, seed=seed, leader_arms_late=late).run()


def test_simulator_detects_single_cta_release():
    """negative control: a commit that releases the stage only in the leader lets ... nothing; releasing it only in the PEER lets
    the leader's producer starve (deadlock), and arming the full barrier with one CTA's bytes lets the MMA start too early"""
    class HalfBytes(PairSim):
        def producer(self, cta):
            it = 0
            for tile in range(self.ntiles):
                for kb in range(self.num_kb):
                    s, ph = it % STAGES, (it // STAGES) & 1
                    yield ("wait", self.empty[cta][s], ph ^ 1)
                    if cta == 0:
                        self.full[s].arrive_expect_tx(100)      # only the leader's own bytes
                    self.tma(self.stage[cta][s], (tile, kb), self.full[s], 100)
                    yield ("delay", 0.01)
                    it += 1

    with pytest.raises(AssertionError):
        for seed in range(40):
            HalfBytes(3, 10, n_epi=2, seed=seed).run()
