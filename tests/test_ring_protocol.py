"""Host-only model of the weight ring of the barrier-free FP32 kernel (21cmvae_b200/csrc/fp32_pipe_kernel.cuh): eight consumer
warps, WST slots, stage g + DIST issued by warp (g mod 8) when it OPENS its stage g, after waiting for `empty` of the slot's
previous occupant (stage g + DIST - WST); a warp waits for `full` of its stage, reads, releases with one arrival on `empty`.
The model runs the warps under random (adversarial) interleavings with random copy latencies and checks what the kernel's header
comment claims: no deadlock, a slot is never overwritten before all eight warps have released its previous stage, a warp only ever
reads the stage it expects, and mbarrier PARITY waits (one bit of phase) cannot be fooled by a barrier running a phase ahead.
"""
import random

import pytest

NW = 8


class Ring:
    def __init__(self, wst, dist, total, rng, latency):
        self.wst, self.dist, self.total, self.rng, self.latency = wst, dist, total, rng, latency
        self.full_phase = [0] * wst       # completed phases of full[slot]
        self.empty_arrivals = [0] * wst   # arrivals in the current phase of empty[slot]
        self.empty_phase = [0] * wst      # completed phases of empty[slot]
        self.content = [None] * wst       # stage whose bytes the slot holds (None: copy in flight / never filled)
        self.in_flight = []               # (ready_time, stage)
        self.released = {}                # stage -> number of warps that released it
        self.clock = 0

    # mbarrier.try_wait.parity semantics: true iff the phase with that parity bit is complete, i.e. the barrier's current
    # (incomplete) phase has the OTHER parity
    def full_done(self, slot, parity):
        return (self.full_phase[slot] & 1) != parity

    def empty_done(self, slot, parity):
        return (self.empty_phase[slot] & 1) != parity

    def issue(self, stage):
        slot = stage % self.wst
        prev = stage - self.wst
        assert prev < 0 or self.released.get(prev, 0) == NW, f"stage {stage} overwrites stage {prev} before all warps released it"
        self.content[slot] = None
        self.in_flight.append((self.clock + self.rng.randint(1, self.latency), stage))

    def tick(self):
        self.clock += 1
        for item in [x for x in self.in_flight if x[0] <= self.clock]:
            self.in_flight.remove(item)
            slot = item[1] % self.wst
            self.content[slot] = item[1]
            self.full_phase[slot] += 1   # expect_tx arrival + all bytes: phase complete

    def release(self, stage):
        slot = stage % self.wst
        self.released[stage] = self.released.get(stage, 0) + 1
        self.empty_arrivals[slot] += 1
        if self.empty_arrivals[slot] == NW:
            self.empty_arrivals[slot] = 0
            self.empty_phase[slot] += 1


def run(wst, dist, total, seed, latency=40, work=6):
    rng = random.Random(seed)
    ring = Ring(wst, dist, total, rng, latency)
    # per-warp state, exactly the registers of `Pipe`: consumer slot / parity / stage count, producer slot / parity / cursor
    warps = [dict(g=0, cs=0, cph=0, ps=0, pph=1, pidx=0, phase="open", busy=0) for _ in range(NW)]
    for w in warps:  # the DIST stages no consumer iteration issues (warp 0 issues, every warp advances its cursor copy)
        for _ in range(dist):
            if w is warps[0] and w["pidx"] < total:
                assert ring.empty_done(w["ps"], w["pph"])  # fresh barrier: parity-1 wait passes at once
                ring.issue(w["pidx"])
            w["ps"] += 1
            if w["ps"] == wst:
                w["ps"], w["pph"] = 0, w["pph"] ^ 1
            w["pidx"] += 1
    idle = 0
    while any(w["g"] < total for w in warps):
        ring.tick()
        i = rng.randrange(NW)
        w = warps[i]
        progressed = False
        if w["g"] >= total:
            pass
        elif w["phase"] == "open":  # this warp's turn to issue stage g + DIST?
            if w["g"] % NW == i and w["pidx"] < total:
                if ring.empty_done(w["ps"], w["pph"]):
                    ring.issue(w["pidx"])
                    w["phase"] = "advance"
                    progressed = True
            else:
                w["phase"] = "advance"
                progressed = True
        elif w["phase"] == "advance":
            w["ps"] += 1
            if w["ps"] == wst:
                w["ps"], w["pph"] = 0, w["pph"] ^ 1
            w["pidx"] += 1
            w["phase"] = "wait"
            progressed = True
        elif w["phase"] == "wait":
            if ring.full_done(w["cs"], w["cph"]):
                assert ring.content[w["cs"]] == w["g"], f"warp {i} expects stage {w['g']} in slot {w['cs']}, finds {ring.content[w['cs']]}"
                w["phase"], w["busy"] = "compute", rng.randint(1, work)
                progressed = True
        elif w["phase"] == "compute":
            w["busy"] -= 1
            assert ring.content[w["cs"]] == w["g"], "slot overwritten while a warp was reading it"
            if w["busy"] == 0:
                ring.release(w["g"])
                w["cs"] += 1
                if w["cs"] == wst:
                    w["cs"], w["cph"] = 0, w["cph"] ^ 1
                w["g"] += 1
                w["phase"] = "open"
            progressed = True
        idle = 0 if progressed or ring.in_flight else idle + 1
        assert idle < 20000, f"deadlock: {[(x['g'], x['phase']) for x in warps]}"
    assert all(ring.released.get(s, 0) == NW for s in range(total))


@pytest.mark.parametrize("wst,dist", [(6, 3), (4, 2), (3, 2), (3, 1)])
def test_ring_protocol_is_deadlock_free_and_never_overwrites_live_slots(wst, dist):
    for seed in range(25):
        run(wst, dist, total=145 * 2 + 7, seed=seed)


def test_ring_protocol_with_slow_copies_and_fast_warps():
    """Copy latency far above a stage's compute time: the warps queue up behind `full`, nothing else changes."""
    for seed in range(5):
        run(6, 3, total=200, seed=100 + seed, latency=400, work=2)


def test_a_ring_as_deep_as_its_prefetch_distance_is_refused_by_the_model():
    """DIST = WST would let the issue of stage g + WST wait for stage g, which the issuing warp itself still holds."""
    with pytest.raises(AssertionError):
        run(3, 3, total=60, seed=1)
