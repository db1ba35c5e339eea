"""Discrete-event simulation of the barrier protocol of the EXPERIMENTAL persistent windowed-attention kernel
(csrc/attn_tc_v3.cu): work items instead of key tiles, a single Q'K' buffer released right after QK^T, single V^T and P
buffers, two S buffers whose first columns are re-used for O.  Infrastructure shared with test_attn_v2_protocol_sim.py."""
import pytest

from tests.test_attn_v2_protocol_sim import Bar, Buf, Sim


class Sim3(Sim):
    def __init__(self, n_items, n_soft, seed):
        super().__init__(n_items, n_soft, seed)
        self.barQK_full, self.barQK_empty = Bar(1), Bar(1)
        self.barV_full1, self.barV_empty1 = Bar(1), Bar(1)
        self.QK, self.V1 = Buf("QK"), Buf("V")
        self.SO = [Buf("S0/O0"), Buf("S1/O1")]      # S_i and O_i live in the same TMEM columns
        self.stored = [0] * n_items

    def producer(self):
        for i in range(self.ntiles):
            if i >= 1:
                yield ("wait", self.barQK_empty, (i - 1) & 1)
            self.tma(self.QK, i, self.barQK_full)
            if i >= 1:
                yield ("wait", self.barV_empty1, (i - 1) & 1)
            self.tma(self.V1, i, self.barV_full1)
            yield ("delay", self.rng.uniform(0.0, 0.3))

    def issue_S(self, i):
        b, u = i & 1, i >> 1
        yield ("wait", self.barQK_full, i & 1)
        if u >= 1:
            yield ("wait", self.barS_empty[b], (u - 1) & 1)
        self.mma([self.QK], self.SO[b], ("S", i), [self.barS_full[b], self.barQK_empty], rtag=i)

    def mma(self, reads, write, tile, commits, rtag=None):
        self.mma_queue.append((reads, write, tile, commits, rtag))
        self._pump()

    def _pump(self):
        if self.mma_busy or not self.mma_queue:
            return
        reads, write, tile, commits, rtag = self.mma_queue.pop(0)
        self.mma_busy = True
        for b in reads:
            b.read_begin(rtag)
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

    def mma_thread(self):
        yield from self.issue_S(0)
        for i in range(self.ntiles):
            b = i & 1
            if i + 1 < self.ntiles:
                yield from self.issue_S(i + 1)
            yield ("wait", self.barP_full, i & 1)
            yield ("wait", self.barV_full1, i & 1)
            # P.V_i: reads P and V (both tagged i), writes O_i over the columns of S_i
            self.mma([self.P, self.V1], self.SO[b], ("O", i), [self.barO_full, self.barV_empty1], rtag=i)
            yield ("delay", self.rng.uniform(0.0, 0.2))

    def softmax_thread(self, tid):
        for i in range(self.ntiles):
            b, u = i & 1, i >> 1
            yield ("wait", self.barS_full[b], u & 1)
            self.SO[b].read_begin(("S", i))              # max pass + exp pass
            self.p_started[i] = self.p_started.get(i, 0) + 1
            if self.p_started[i] == 1:
                self.P.write_begin(i)
            else:
                assert self.P.writer == i
            yield ("delay", self.rng.uniform(0.4, 2.5))
            self.SO[b].read_end()
            self.p_done[i] = self.p_done.get(i, 0) + 1
            if self.p_done[i] == self.n_soft:
                self.P.write_end(i)
            self.barP_full.arrive()
            yield ("wait", self.barO_full, i & 1)        # epilogue of item i
            self.SO[b].read_begin(("O", i))
            yield ("delay", self.rng.uniform(0.1, 0.8))
            self.SO[b].read_end()
            if tid == 0:
                self.stored[i] += 1
                self.o_accumulated[i] += 1
            self.barS_empty[b].arrive()


@pytest.mark.parametrize("n_items", [1, 2, 3, 8, 55])
def test_v3_protocol_terminates_without_hazards(n_items):
    for seed in range(50):
        Sim3(n_items, n_soft=3, seed=seed).run()


def test_v3_simulator_detects_early_release_of_the_s_columns():
    """negative control: releasing the S/O columns when P is published (instead of after the epilogue read O) lets S_{i+2}
    overwrite O_i before it has been stored"""
    class EarlyRelease(Sim3):
        def softmax_thread(self, tid):
            for i in range(self.ntiles):
                b, u = i & 1, i >> 1
                yield ("wait", self.barS_full[b], u & 1)
                self.SO[b].read_begin(("S", i))
                self.p_started[i] = self.p_started.get(i, 0) + 1
                if self.p_started[i] == 1:
                    self.P.write_begin(i)
                yield ("delay", self.rng.uniform(0.4, 2.5))
                self.SO[b].read_end()
                self.p_done[i] = self.p_done.get(i, 0) + 1
                if self.p_done[i] == self.n_soft:
                    self.P.write_end(i)
                self.barP_full.arrive()
                self.barS_empty[b].arrive()              # too early
                yield ("wait", self.barO_full, i & 1)
                yield ("delay", self.rng.uniform(0.5, 6.0))
                self.SO[b].read_begin(("O", i))
                yield ("delay", 0.2)
                self.SO[b].read_end()
                if tid == 0:
                    self.o_accumulated[i] += 1

    with pytest.raises(AssertionError):
        for seed in range(80):
            EarlyRelease(9, n_soft=2, seed=seed).run()
