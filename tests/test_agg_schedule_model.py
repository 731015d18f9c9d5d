"""Host-side model of the aggregation kernel's single-sweep schedule (csrc/aggregate_tc.cu: `tc_item`, the epilogue's
front / back halves with the pending item, the MMA issuer's TMEM buffer ring, `agg_tc_run`'s choice of J and grid).

The kernel lets sibling CTAs wait for each other inside a persistent grid, and an epilogue that holds TMEM buffers across
that wait.  DESIGN.md 4.1 claims this cannot deadlock for any mix of empty / single-sweep / two-sweep items, any J and an
unaligned grid; this discrete simulation of the wait-for structure checks the claim (and that every accumulator buffer is
handed back exactly once) on randomized launches, with a negative control that the model does detect a deadlock.
No GPU and no oracle involved: it is a test of the schedule's logic, the CUDA code mirrors it line by line.
"""
import random

import pytest

K_BUFS = 4          # kTcBufs: 512 TMEM columns / 128-column passes
SUB_CHUNKS = 2      # kTcSubChunks
EMPTY, RES, LONG = 0, 1, 2


def choose_j(P, n_rest, num_sms, resident=True, look_ahead=False):
    """agg_tc_run: J = ceil(P / passes per CTA), doubled while there are few items; grid = min(items, SMs)."""
    per_cta = K_BUFS // 2 if look_ahead else K_BUFS
    jres = -(-P // per_cta)
    resident = resident and jres <= 32 and 2 * jres <= num_sms
    J = jres if resident else 1
    while n_rest * J < 2 * num_sms and J * 2 <= P:
        J *= 2
    n_items = n_rest * J
    return J, n_items, min(n_items, num_sms), resident


class Launch:
    def __init__(self, P, J, grid, kinds, nsub_long, resident, force_la=None):
        self.P, self.J, self.G, self.kinds, self.nsub_long, self.resident = P, J, grid, kinds, nsub_long, resident
        self.n_items = len(kinds) * J
        self.la = resident and J > 1 and 2 * (-(-P // J)) <= K_BUFS
        if force_la is not None:
            self.la = force_la
        self.posted = [0] * len(kinds)
        self.mma_done = [set() for _ in range(grid)]
        self.released = [dict() for _ in range(grid)]

    def item(self, idx):
        j, rest = idx % self.J, idx // self.J
        kind = self.kinds[rest]
        np_ = (j + 1) * self.P // self.J - j * self.P // self.J
        # tc_item: single sweep needs one accumulator chain per pass and <= kTcBufs passes per sibling
        res = self.resident and kind == RES and -(-self.P // self.J) <= K_BUFS
        if kind == RES and not res:
            kind = LONG                                        # the two-sweep path with one chain per pass
            nsub = 1
        else:
            nsub = self.nsub_long if kind == LONG else 1
        return kind, res, np_, nsub

    # --- MMA issuer: uses the TMEM buffers as a ring, one use per (pass, chain) -------------------------------------------
    def mma_thread(self, c):
        u = 0
        for idx in range(c, self.n_items, self.G):
            kind, res, np_, nsub = self.item(idx)
            if kind == EMPTY:
                continue
            for _ in range(((0 if res else self.P) + np_) * nsub):
                if u >= K_BUFS:
                    yield ("released", c, u - K_BUFS)          # bar_tempty of the buffer's previous use
                self.mma_done[c].add(u)
                u += 1

    def release(self, c, u):
        assert u in self.mma_done[c], "buffer handed back before its accumulator was complete"
        self.released[c][u] = self.released[c].get(u, 0) + 1

    # --- epilogue: front half (sums of squares, post), back half (wait for the siblings, write) -----------------------------
    def epilogue_thread(self, c):
        ti, pend, idx = 0, None, c
        while idx < self.n_items or pend is not None:
            have = idx < self.n_items
            cur = self.item(idx) if have else None
            front = have and not (cur[0] != EMPTY and pend is not None and not (cur[1] and self.la))
            cur_ti0 = ti
            if front:
                kind, res, np_, nsub = cur
                if kind == EMPTY:
                    idx += self.G
                    continue
                for _ in range((np_ if res else self.P) * nsub):
                    yield ("mma", c, ti)
                    if not res:
                        self.release(c, ti)
                    ti += 1
                if res and self.J > 1:
                    self.posted[idx // self.J] += 1
            for w in (0, 1):
                if w == 0:
                    if pend is None:
                        continue
                    (wid, ti0, it), pend = pend, None
                else:
                    if not front:
                        continue
                    if cur[1] and self.la:
                        pend = (idx, cur_ti0, cur)
                        continue
                    wid, ti0, it = idx, cur_ti0, cur
                kind, res, np_, nsub = it
                if res and self.J > 1:
                    yield ("siblings", wid // self.J)
                tw = ti0 if res else ti
                for _ in range(np_ * nsub):
                    if not res:
                        yield ("mma", c, tw)
                    self.release(c, tw)
                    tw += 1
                if not res:
                    ti = tw
            if front:
                idx += self.G

    def ready(self, cond):
        if cond[0] == "released":
            return self.released[cond[1]].get(cond[2], 0) > 0
        if cond[0] == "mma":
            return cond[2] in self.mma_done[cond[1]]
        return self.posted[cond[1]] == self.J

    def run(self):
        """Round-robin over all threads of all (resident) CTAs; returns True when every thread finished."""
        threads = []
        for c in range(self.G):
            threads.append([self.mma_thread(c), None, False])
            threads.append([self.epilogue_thread(c), None, False])
        progress = True
        while progress:
            progress = False
            for t in threads:
                if t[2]:
                    continue
                while True:
                    if t[1] is not None and not self.ready(t[1]):
                        break
                    try:
                        t[1] = next(t[0])
                        progress = True
                    except StopIteration:
                        t[2] = True
                        progress = True
                        break
        return all(t[2] for t in threads)

    def check_buffers(self):
        for c in range(self.G):
            assert set(self.released[c]) == self.mma_done[c]
            assert all(v == 1 for v in self.released[c].values()), "a TMEM buffer was handed back twice"


def _kinds(rng, n_rest, p_empty, p_long):
    return [EMPTY if (x := rng.random()) < p_empty else (LONG if x < p_empty + p_long else RES) for _ in range(n_rest)]


@pytest.mark.parametrize("P", [1, 2, 4, 6, 12, 13])
@pytest.mark.parametrize("look_ahead", [False, True])
def test_single_sweep_schedule_cannot_deadlock(P, look_ahead):
    rng = random.Random(1000 * P + look_ahead)
    for trial in range(12):
        num_sms = rng.choice([4, 7, 16, 148])
        n_rest = rng.choice([1, 2, 3, 17, 64, 300])
        J, n_items, grid, resident = choose_j(P, n_rest, num_sms, True, look_ahead)
        kinds = _kinds(rng, n_rest, rng.choice([0.0, 0.1, 0.5]), rng.choice([0.0, 0.05, 0.4]))
        L = Launch(P, J, grid, kinds, nsub_long=rng.choice([1, 2, 3]), resident=resident)
        assert L.n_items == n_items
        assert L.run(), f"deadlock: P={P} J={J} grid={grid} n_rest={n_rest} la={L.la}"
        L.check_buffers()
        assert all(L.posted[r] in (0, J) for r in range(n_rest))


def test_two_sweep_schedule_and_channel_split():
    # SEGVLAD_AGG_RESIDENT=0: no item waits for another CTA, whatever the split
    rng = random.Random(5)
    for P in (4, 12):
        J, n_items, grid, resident = choose_j(P, 5, 148, resident=False)
        assert not resident and J > 1                         # few items: the write sweep is split over channels
        L = Launch(P, J, grid, _kinds(rng, 5, 0.2, 0.3), nsub_long=2, resident=False)
        assert L.run() and sum(L.posted) == 0
        L.check_buffers()


def test_default_bench_shape():
    # 16 images x 128 SuperSegments, K = 64, D_t = 1536: 1024 (tile, cluster) blocks, 12 passes -> J = 3 siblings of 4 passes
    J, n_items, grid, resident = choose_j(12, 16 * 64, 148)
    assert (J, n_items, grid, resident) == (3, 3072, 148, True)
    L = Launch(12, J, grid, [RES] * 1024, 1, resident)
    assert not L.la and L.run()
    J2, _, _, _ = choose_j(12, 16 * 64, 148, look_ahead=True)
    assert J2 == 6 and Launch(12, J2, 148, [RES] * 64, 1, True).la


def test_model_detects_a_deadlock():
    # negative control: a look-ahead with items that fill the whole TMEM holds every buffer while it waits for the next
    # item's accumulators -- the kernel only enables the look-ahead when two items fit (2 * ceil(P / J) <= kTcBufs)
    L = Launch(P=12, J=3, grid=6, kinds=[RES] * 8, nsub_long=1, resident=True, force_la=True)
    assert not L.run()
