"""Protocol model of the peer-memory exchange (simple_dqn_b200/csrc/comm_p2p.cuh), run on CPU.

The CUDA kernels are validated on hardware (tests/test_gpu_multi.py, tools/mgpu_check.py); this file checks the
PROTOCOL they implement — which buffer a rank may write when, and what a flag word may be compared with — under
arbitrary interleavings of W ranks, which a 2-GPU box cannot enumerate:

* k_xll (one-shot LL all-reduce): every line carries {data, epoch}; receive areas are double-buffered by the parity
  of the layer's epoch; a receiver accepts a line only when its flag EQUALS the epoch it expects; the layer is the
  channel (a line's flags only ever carry its own layer's epoch sequence).
* k_xpush / k_xwait (plane gather of H3 for fc1_wgrad): gather areas double-buffered by the parity of the push epoch,
  one monotonic flag per (channel, source); the consumer waits for flag >= its own epoch.
* k_xgather_ll (round 2: the dZ4 rows): the same LL lines as k_xll on a channel of their own, but the receiver STORES
  each source's payload instead of adding it; its known-answer test at comm_init runs on the same channel, so the
  epoch sequence of a line never repeats.

Ranks are generators that yield before every shared-memory access; a seeded scheduler picks who moves next.  A rank
finishes all exchanges of step s before it starts step s + 1 (stream order of the captured step), and the exchanges
of one step run in any order (they sit on different graph branches).
"""
import random

import pytest

LAYERS = (0, 1, 2, 4)          # LL channels (fc1 = 3 is gathered, not reduced)
DZ = 6                         # LL channel of the dZ4 all-gather
PARITY_MASK = [1]              # 1: receive areas double-buffered by epoch parity (the design); 0: single buffer


class World:
    def __init__(self, W):
        self.W = W
        # LL receive areas: [rank][parity][src][layer] -> (payload, flag)
        self.ll = [[[{l: (None, 0) for l in LAYERS + (DZ,)} for _ in range(W)] for _ in range(2)] for _ in range(W)]
        # gather areas: [rank][chan][parity][src] -> payload; push flags: [rank][chan][src]
        self.gat = [[[[None] * W for _ in range(2)] for _ in range(2)] for _ in range(W)]
        self.pflag = [[[0] * W for _ in range(2)] for _ in range(W)]


def ll_exchange(w, r, layer, epoch_of, payload, chan_of=lambda l: l):
    """One k_xll call of rank r for `layer`; epoch_of[chan] is this rank's device-side counter."""
    chan = chan_of(layer)
    e = epoch_of[chan] + 1
    par = e & PARITY_MASK[0]
    for p in range(w.W):                       # push to every peer (never blocks)
        if p != r:
            yield
            w.ll[p][par][r][layer] = (payload, e)
    got = []
    for p in range(w.W):                       # collect in rank order
        if p == r:
            got.append(payload)
            continue
        while True:
            yield
            val, flag = w.ll[r][par][p][layer]
            if flag == e:
                got.append(val)
                break
    epoch_of[chan] = e
    return got


def push(w, r, chan, epoch_of, payload):
    e = epoch_of[chan] + 1
    par = e & 1
    for p in range(w.W):
        yield
        w.gat[p][chan][par][r] = payload
    for p in range(w.W):                       # flags after the data (release)
        yield
        w.pflag[p][chan][r] = e
    epoch_of[chan] = e


def wait_and_read(w, r, epoch_of, chans=(0, 1)):
    out = []
    for chan in chans:
        e = epoch_of[chan]
        for p in range(w.W):
            while True:
                yield
                if w.pflag[r][chan][p] >= e:
                    break
        out.append([w.gat[r][chan][e & 1][p] for p in range(w.W)])
    return out


def rank_program(w, r, steps, rng, kat_chan_of, errors):
    ll_epoch, push_epoch = {}, {0: 0, 1: 0}
    for c in set(LAYERS) | {5, DZ}:
        ll_epoch[c] = 0
    # comm_init known-answer tests
    for l in LAYERS:
        got = yield from ll_exchange(w, r, l, ll_epoch, ("kat", l), chan_of=kat_chan_of)
        if got != [("kat", l)] * w.W:
            errors.append(("kat", r, l, got))
    yield from push(w, r, 0, push_epoch, ("kat", r))
    yield from push(w, r, 1, push_epoch, ("kat", r))
    yield from wait_and_read(w, r, push_epoch)
    got = yield from ll_exchange(w, r, DZ, ll_epoch, ("katdz", r))
    if got != [("katdz", p) for p in range(w.W)]:
        errors.append(("kat-dz", r, got))
    for s in range(1, steps + 1):
        # H3 push early in the step, dZ4 push after the head, then the consumer; LL layers in any order
        yield from push(w, r, 0, push_epoch, ("h3", s, r))
        order = list(LAYERS)
        rng.shuffle(order)
        pending = [ll_exchange(w, r, l, ll_epoch, ("g", s, l, r)) for l in order]

        def fc1_branch():          # dZ4 all-gather in LL lines, H3 flags polled by the same kernel
            dz = yield from ll_exchange(w, r, DZ, ll_epoch, ("dz4", s, r))
            (h3,) = yield from wait_and_read(w, r, push_epoch, chans=(0,))
            return h3, dz
        pending.append(fc1_branch())
        results = {}
        live = list(range(len(pending)))
        while live:                            # branches of one step advance independently
            i = rng.choice(live)
            try:
                next(pending[i])
                yield
            except StopIteration as done:
                results[i] = done.value
                live.remove(i)
        for i, l in enumerate(order):
            want = [("g", s, l, p) for p in range(w.W)]
            if results[i] != want:
                errors.append(("ll", r, s, l, results[i]))
        h3, dz = results[len(order)]
        if h3 != [("h3", s, p) for p in range(w.W)] or dz != [("dz4", s, p) for p in range(w.W)]:
            errors.append(("gather", r, s, h3, dz))


def simulate(W, steps, seed, kat_chan_of=lambda l: l, max_ticks=2_000_000):
    w = World(W)
    rng = random.Random(seed)
    errors = []
    progs = {r: rank_program(w, r, steps, random.Random(seed * 131 + r), kat_chan_of, errors) for r in range(W)}
    ticks = 0
    while progs:
        # skewed scheduling: now and then one rank sprints, another stalls
        r = rng.choice(list(progs))
        burst = rng.choice((1, 1, 1, 5, 40))
        for _ in range(burst):
            try:
                next(progs[r])
            except StopIteration:
                del progs[r]
                break
        ticks += burst
        assert ticks < max_ticks, "protocol dead-locked (or the model spins)"
    return errors


@pytest.mark.parametrize("W", [2, 3, 4, 8])
def test_exchange_protocol_delivers_every_step_exactly(W):
    for seed in range(20):
        assert simulate(W, steps=8, seed=seed) == []


def test_model_shows_why_the_ll_areas_are_double_buffered():
    """With one receive area a rank that runs ahead overwrites lines its peer has not read yet: the reader then sees
    a flag from the future and never accepts it (dead-lock) or, with >=, would sum the wrong step."""
    PARITY_MASK[0] = 0
    try:
        broken = 0
        for seed in range(12):
            try:
                broken += bool(simulate(3, steps=4, seed=seed, max_ticks=200_000))
            except AssertionError:
                broken += 1
        assert broken > 0
    finally:
        PARITY_MASK[0] = 1


def test_model_catches_a_shared_kat_channel():
    """The bug found on hardware in round 1: known-answer exchanges run on a separate channel left lines whose flag
    equalled the first epoch of the per-layer channels, so a fast reader accepted stale data.  The model must see it."""
    bad = [simulate(2, steps=2, seed=s, kat_chan_of=lambda l: 5) for s in range(12)]
    assert any(bad), "a shared KAT channel went unnoticed"
