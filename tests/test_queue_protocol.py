"""The two-ended work queue of ctr_queue8_kernel / ctr_queue_kernel / xts_sectors_hybrid_kernel / ecb_dec_hybrid_kernel (csrc/uaes_kernels.cu, q_post /
q_front / q_back), restated in Python and checked on random interleavings: table-driven warps claim units from
the front, bitsliced warps from the back, through ONE atomic add on a packed (front, back) word; a claim is
valid iff front + back < units *at its own place in the atomic order*.

Invariant under every interleaving: each unit is served by exactly one claimant, none is left over, and
every warp stops after exactly one failed claim.  (This is host-side logic: no GPU, no oracle.)  The same
file pins the unit geometry the launcher derives from a counter range (uaes_kernels.cu, launch_ctr_nr).
"""
import random

import pytest

KQ_NONE = None


class Queue:
    """the packed 64-bit word: front in the low 32 bits, back in the high 32 bits"""

    def __init__(self, units):
        self.word, self.units = 0, units

    def post(self, inc):
        old = self.word
        self.word = (self.word + inc) & ((1 << 64) - 1)      # atom.global.add.u64
        return old

    def front(self, old):
        f, b = old & 0xFFFFFFFF, old >> 32
        return f if f + b < self.units else KQ_NONE

    def back(self, old):
        f, b = old & 0xFFFFFFFF, old >> 32
        return self.units - 1 - b if f + b < self.units else KQ_NONE


def run(units, n_front, n_back, rng):
    """every warp: claim, work, claim ... with the NEXT claim posted before the current unit is worked on
    (the kernels claim a unit ahead); the scheduler picks a random runnable warp at each step"""
    q = Queue(units)
    served = {}
    warps = [("f", i) for i in range(n_front)] + [("b", i) for i in range(n_back)]
    state = {w: {"cur": None, "posted": None, "done": False, "failed": 0} for w in warps}
    for w in warps:                                           # first claim: posted and read at once
        st = state[w]
        old = q.post(1 if w[0] == "f" else 1 << 32)
        st["cur"] = q.front(old) if w[0] == "f" else q.back(old)
        if st["cur"] is KQ_NONE:
            st["done"], st["failed"] = True, 1
    while not all(s["done"] for s in state.values()):
        w = rng.choice([w for w in warps if not state[w]["done"]])
        st = state[w]
        if st["posted"] is None:                              # top of a unit: post the next claim, then work
            st["posted"] = q.post(1 if w[0] == "f" else 1 << 32)
            assert st["cur"] not in served, f"unit {st['cur']} served twice"
            served[st["cur"]] = w
        else:                                                 # end of the unit: read the answer
            nxt = q.front(st["posted"]) if w[0] == "f" else q.back(st["posted"])
            st["posted"] = None
            if nxt is KQ_NONE:
                st["done"], st["failed"] = True, st["failed"] + 1
            st["cur"] = nxt
    return served, state


@pytest.mark.parametrize("units,n_front,n_back", [(0, 3, 2), (1, 4, 4), (7, 12, 4), (100, 12, 4), (1000, 37, 11),
                                                  (5, 0, 4), (5, 6, 0), (64, 1, 1)])
def test_every_unit_is_served_exactly_once(units, n_front, n_back):
    for seed in range(40):
        served, state = run(units, n_front, n_back, random.Random(seed * 7919 + units))
        assert sorted(served) == list(range(units)), (units, n_front, n_back, seed)
        assert all(s["failed"] == 1 for s in state.values())          # one failed claim ends a warp
        # front claimants hold a prefix, back claimants a suffix: the border is one point
        f_units = sorted(u for u, w in served.items() if w[0] == "f")
        b_units = sorted(u for u, w in served.items() if w[0] == "b")
        assert f_units == list(range(len(f_units))) and b_units == list(range(units - len(b_units), units))


def unit_plan(v0, nblocks, shift=11):
    """launch_ctr_nr: units of 2^shift counters aligned in counter space covering [v0, v0 + nblocks)"""
    unit = 1 << shift
    u0 = v0 & ~(unit - 1)
    return u0, (v0 - u0 + nblocks + unit - 1) >> shift


def test_unit_geometry_covers_the_range_exactly():
    rng = random.Random(5)
    for _ in range(2000):
        v0 = rng.choice([0, 1, 2047, 2048, (1 << 32) - 3, (1 << 56) - 5000, rng.randrange(1 << 56)])
        n = rng.choice([1, 2047, 2048, 2049, 4096, rng.randrange(1, 1 << 22)])
        u0, units = unit_plan(v0, n)
        assert u0 <= v0 < u0 + 2048 and u0 % 2048 == 0
        assert u0 + units * 2048 >= v0 + n > u0 + (units - 1) * 2048          # the last unit is not empty
        # blocks of unit u clipped to the call (what both kinds of warps do with k = counter - v0) add up to the call
        covered = sum(max(0, min(v0 + n, u0 + (u + 1) * 2048) - max(v0, u0 + u * 2048)) for u in range(units))
        assert covered == n
