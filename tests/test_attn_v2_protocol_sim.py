"""Discrete-event simulation of the barrier protocol of the EXPERIMENTAL attn_tc_v2_kernel (csrc/attn_tc_v2.cu).

A warp-specialised kernel with a wrong hand-over either hangs (a strike on the GPU box) or races silently.  The three roles of
the kernel are restated here as coroutines over simulated mbarriers (phase / parity semantics of `mbarrier.try_wait.parity`),
an in-order tensor pipe with `tcgen05.commit` markers and an asynchronous TMA engine, with randomised latencies.  Every run
must terminate (no deadlock, no phase aliasing) and must never overlap a write to a buffer with a read of its previous
contents (K', V stages, the single P buffer, the two S buffers, the single O buffer).  The coroutines mirror the kernel
line by line: a change to the kernel's protocol must be made here first.

Finding of the simulation: barS_empty is implied by the other hand-overs (S_{j+2} is issued only after P_j was published, i.e.
after every softmax thread finished reading S_j); it is kept in the kernel as a cheap safety net."""
import heapq
import random

import pytest


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier expects in one phase"
        if self.pending == 0:
            self.pending = self.count
            self.phase += 1


class Buf:
    """tracks who reads / writes a buffer to detect overlapping use of different tiles"""

    def __init__(self, name):
        self.name, self.readers, self.writer, self.content = name, 0, None, None

    def write_begin(self, tile):
        assert self.readers == 0 and self.writer is None, f"{self.name}: write of tile {tile} while busy (content {self.content})"
        self.writer = tile

    def write_end(self, tile):
        assert self.writer == tile
        self.writer, self.content = None, tile

    def read_begin(self, tile):
        assert self.writer is None and self.content == tile, f"{self.name}: read of tile {tile} sees {self.content} / writer {self.writer}"
        self.readers += 1

    def read_end(self):
        self.readers -= 1


class Sim:
    def __init__(self, ntiles, n_soft, seed):
        self.rng = random.Random(seed)
        self.now, self.events, self.seq = 0.0, [], 0
        self.ntiles, self.n_soft = ntiles, n_soft
        self.barQ, self.barK_full, self.barK_empty = Bar(1), Bar(1), Bar(1)
        self.barV_full, self.barV_empty = [Bar(1), Bar(1)], [Bar(1), Bar(1)]
        self.barS_full, self.barS_empty = [Bar(1), Bar(1)], [Bar(n_soft), Bar(n_soft)]
        self.barP_full, self.barO_full = Bar(n_soft), Bar(1)
        self.K, self.V = Buf("K"), [Buf("V0"), Buf("V1")]
        self.S, self.P, self.O = [Buf("S0"), Buf("S1")], Buf("P"), Buf("O")
        self.mma_queue, self.mma_busy = [], False
        self.p_started, self.p_done = {}, {}
        self.o_accumulated = [0] * ntiles

    def at(self, dt, fn):
        self.seq += 1
        heapq.heappush(self.events, (self.now + dt, self.seq, fn))

    # ---- engines
    def tma(self, buf, tile, bar):
        buf.write_begin(tile)

        def done():
            buf.write_end(tile)
            bar.arrive()
        self.at(self.rng.uniform(0.2, 3.0), done)

    def mma(self, reads, write, tile, commits):
        """in-order tensor pipe: op starts when the previous one finished; `commits` arrive when it (and all before) is done"""
        self.mma_queue.append((reads, write, tile, commits))
        self._pump()

    def _pump(self):
        if self.mma_busy or not self.mma_queue:
            return
        reads, write, tile, commits = self.mma_queue.pop(0)
        self.mma_busy = True
        for b in reads:
            b.read_begin(tile)
        write.write_begin(tile)

        def done():
            for b in reads:
                b.read_end()
            write.write_end(tile)
            for bar in commits:
                bar.arrive()
            self.mma_busy = False
            self._pump()
        self.at(self.rng.uniform(0.1, 1.5), done)

    # ---- roles (generators yield ("wait", bar, parity) or ("delay", t))
    def producer(self):
        self.tma(Buf("Q"), 0, self.barQ)
        for j in range(self.ntiles):
            b, u = j & 1, j >> 1
            if j >= 1:
                yield ("wait", self.barK_empty, (j - 1) & 1)
            self.tma(self.K, j, self.barK_full)
            if u >= 1:
                yield ("wait", self.barV_empty[b], (u - 1) & 1)
            self.tma(self.V[b], j, self.barV_full[b])
            yield ("delay", self.rng.uniform(0.0, 0.3))

    def issue_S(self, j):
        b, u = j & 1, j >> 1
        yield ("wait", self.barK_full, j & 1)
        if u >= 1:
            yield ("wait", self.barS_empty[b], (u - 1) & 1)
        self.mma([self.K], self.S[b], j, [self.barS_full[b], self.barK_empty])

    def mma_thread(self):
        yield ("wait", self.barQ, 0)
        yield from self.issue_S(0)
        for j in range(self.ntiles):
            b, u = j & 1, j >> 1
            if j + 1 < self.ntiles:
                yield from self.issue_S(j + 1)
            yield ("wait", self.barP_full, j & 1)
            yield ("wait", self.barV_full[b], u & 1)
            self.mma([self.P, self.V[b]], self.O, j, [self.barO_full, self.barV_empty[b]])
            yield ("delay", self.rng.uniform(0.0, 0.2))

    def softmax_thread(self, tid):
        def accumulate(t):
            yield ("wait", self.barO_full, t & 1)
            self.O.read_begin(t)
            yield ("delay", self.rng.uniform(0.1, 0.6))
            self.O.read_end()
            if tid == 0:
                self.o_accumulated[t] += 1

        for j in range(self.ntiles):
            b, u = j & 1, j >> 1
            yield ("wait", self.barS_full[b], u & 1)
            self.S[b].read_begin(j)                      # max pass
            yield ("delay", self.rng.uniform(0.2, 1.2))
            self.S[b].read_end()
            if j >= 1:
                yield from accumulate(j - 1)
            self.S[b].read_begin(j)                      # exp pass + P write (every softmax thread writes its own rows)
            self.p_started[j] = self.p_started.get(j, 0) + 1
            if self.p_started[j] == 1:
                self.P.write_begin(j)                    # asserts that P.V of the previous tile is not reading P any more
            else:
                assert self.P.writer == j, f"P: rows of tile {j} written while the buffer belongs to {self.P.writer}/{self.P.content}"
            yield ("delay", self.rng.uniform(0.2, 1.5))
            self.S[b].read_end()
            self.p_done[j] = self.p_done.get(j, 0) + 1
            if self.p_done[j] == self.n_soft:
                self.P.write_end(j)
            self.barS_empty[b].arrive()
            self.barP_full.arrive()
        yield from accumulate(self.ntiles - 1)

    # ---- scheduler
    def run(self):
        agents = [self.producer(), self.mma_thread()] + [self.softmax_thread(t) for t in range(self.n_soft)]
        blocked = {}  # agent index -> (bar, parity)
        alive = set(range(len(agents)))

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
            assert guard < 200000, "simulation does not terminate"
            progressed = False
            for i, (bar, parity) in list(blocked.items()):
                # mbarrier.try_wait.parity: succeeds once the phase with this parity has completed
                if (bar.phase & 1) != parity:
                    # phase aliasing check: the waiter must not be more than one phase behind
                    del blocked[i]
                    step(i)
                    progressed = True
            if progressed:
                continue
            if not self.events:
                raise AssertionError(f"deadlock at t={self.now:.2f}: blocked agents {sorted(blocked)}")
            self.now, _, fn = heapq.heappop(self.events)
            fn()
        assert not self.mma_queue and not self.mma_busy or True
        # drain outstanding async work
        while self.events:
            self.now, _, fn = heapq.heappop(self.events)
            fn()
        assert self.o_accumulated == [1] * self.ntiles


@pytest.mark.parametrize("ntiles", [2, 3, 4, 7, 32])
def test_protocol_terminates_without_hazards(ntiles):
    for seed in range(60):
        Sim(ntiles, n_soft=3, seed=seed).run()


def test_simulator_detects_a_broken_protocol():
    """sanity of the simulator itself: dropping the K'-buffer hand-over must be caught as a hazard or a deadlock"""
    class Broken(Sim):
        def producer(self):
            self.tma(Buf("Q"), 0, self.barQ)
            for j in range(self.ntiles):
                b, u = j & 1, j >> 1
                self.tma(self.K, j, self.barK_full)      # no wait on barK_empty: overwrites K' while QK^T may still read it
                if u >= 1:
                    yield ("wait", self.barV_empty[b], (u - 1) & 1)
                self.tma(self.V[b], j, self.barV_full[b])
                yield ("delay", 0.01)

    with pytest.raises(AssertionError):
        for seed in range(40):
            Broken(6, n_soft=2, seed=seed).run()


def test_simulator_detects_missing_v_empty_and_early_p_write():
    class NoVEmpty(Sim):
        def producer(self):
            self.tma(Buf("Q"), 0, self.barQ)
            for j in range(self.ntiles):
                b = j & 1
                if j >= 1:
                    yield ("wait", self.barK_empty, (j - 1) & 1)
                self.tma(self.K, j, self.barK_full)
                self.tma(self.V[b], j, self.barV_full[b])   # no wait on barV_empty: P.V of tile j-2 may still read this stage
                yield ("delay", 0.01)

    class EarlyP(Sim):
        """the exp pass (which overwrites the single P buffer) moved BEFORE the deferred accumulation of tile j-1"""

        def softmax_thread(self, tid):
            for j in range(self.ntiles):
                b, u = j & 1, j >> 1
                yield ("wait", self.barS_full[b], u & 1)
                self.S[b].read_begin(j)
                self.p_started[j] = self.p_started.get(j, 0) + 1
                if self.p_started[j] == 1:
                    self.P.write_begin(j)                # P.V of tile j-1 has not been observed complete yet
                yield ("delay", self.rng.uniform(0.2, 1.5))
                self.S[b].read_end()
                self.p_done[j] = self.p_done.get(j, 0) + 1
                if self.p_done[j] == self.n_soft:
                    self.P.write_end(j)
                if j >= 1:
                    yield ("wait", self.barO_full, (j - 1) & 1)
                    self.O.read_begin(j - 1)
                    yield ("delay", 0.1)
                    self.O.read_end()
                    if tid == 0:
                        self.o_accumulated[j - 1] += 1
                self.barS_empty[b].arrive()
                self.barP_full.arrive()
            yield ("wait", self.barO_full, (self.ntiles - 1) & 1)
            if tid == 0:
                self.o_accumulated[self.ntiles - 1] += 1

    for cls in (NoVEmpty, EarlyP):
        with pytest.raises(AssertionError):
            for seed in range(80):
                cls(8, n_soft=2, seed=seed).run()
