"""GPU parity of the replay half: device sampler + TMA gather + state window vs the oracle and
vs golden vectors the reference itself produced.  Bit-exact everywhere (integer/byte work)."""
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN
from helpers import crc, make_args
from oracle.mt19937 import MT19937
from oracle.replay_oracle import (ReplayOracle, StateBufferOracle, decode_frame_tag, indexed_episode_stream,
                                  synthetic_ring)

pytestmark = pytest.mark.gpu


def _mem(size, batch, **kw):
    from simple_dqn_b200 import ReplayMemory
    return ReplayMemory(size, make_args(batch_size=batch), **kw)


def test_golden_cases_from_reference():
    g = np.load(os.path.join(GOLDEN, "replay_golden.npz"))
    for case in g["names"]:
        size, steps, batch, seed, nmb = [int(x) for x in g[case + "/cfg"]]
        tp = float(g[case + "/terminal_p"])
        mem = _mem(size, batch, rng="python")
        for (a, r, s, t) in indexed_episode_stream(steps, seed=seed, terminal_p=tp):
            mem.add(a, r, s, t)                       # exercises the per-step add() path
        assert [mem.count, mem.current] == list(g[case + "/count_current"])
        random.setstate((3, tuple(int(x) for x in g[case + "/mt_before"]), None))
        for i in range(nmb):
            pre, a, r, post, t = mem.getMinibatch()
            assert (mem.last_indexes == g[case + "/indexes"][i]).all(), case
            assert crc(pre) == g[case + "/pre_crc"][i] and crc(post) == g[case + "/post_crc"][i], case
            assert (a == g[case + "/actions"][i]).all() and a.dtype == np.uint8
            assert (r == g[case + "/rewards"][i]).all() and r.dtype == np.int64
            assert (t == g[case + "/terminals"][i]).all() and t.dtype == np.bool_
        assert list(random.getstate()[1]) == [int(x) for x in g[case + "/mt_after"]], case
        assert (pre[0] == g[case + "/last_pre_sample0"]).all()
        assert (mem.getState(-1) == g[case + "/getState_m1"]).all()
        assert (mem.getState(2) == g[case + "/getState_2"]).all()
        # aliasing quirk of the reference: the same persistent buffers come back
        assert pre is mem.prestates and post is mem.poststates


@pytest.mark.parametrize("batch", [32, 256])
def test_sampler_and_gather_match_oracle_midsize(batch):
    size = 20000
    orc = ReplayOracle(size, batch_size=batch)
    synthetic_ring(orc, seed=5, block=500, terminal_p=0.01, count=size - 1234, current=777)
    mem = _mem(size, batch, rng="python")
    mem.add_batch(orc.actions, orc.rewards, orc.screens, orc.terminals)
    mem.set_cursor(orc.count, orc.current)
    random.seed(2024)
    rng = MT19937.from_python(random)
    for _ in range(40):
        before = rng.words_drawn
        op, oa, orr, oq, ot = orc.getMinibatch(rng)
        gp, ga, gr, gq, gt = mem.getMinibatch()
        assert (gp == op).all() and (gq == oq).all()
        assert (ga == oa).all() and (gr == orr).all() and (gt == ot).all()
        assert mem.last_words_consumed == rng.words_drawn - before
    assert list(random.getstate()[1]) == rng.state625()       # host stream stayed in lock-step


def test_device_resident_stream_equals_python_stream():
    size, batch = 5000, 32
    orc = ReplayOracle(size, batch_size=batch)
    synthetic_ring(orc, seed=9, block=100, terminal_p=0.02)
    mem = _mem(size, batch, rng="device")
    mem.add_batch(orc.actions, orc.rewards, orc.screens, orc.terminals)
    mem.set_cursor(orc.count, orc.current)
    random.seed(31337)
    rng = MT19937.from_python(random)
    host_state = random.getstate()
    for _ in range(60):                                       # > 624 words: crosses several twists
        idx = orc.sample_indexes(rng)
        mem.getMinibatch()
        assert (mem.last_indexes == idx).all()
    assert random.getstate() == host_state                    # device mode never touches host random
    assert list(mem.read_device_rng()) == rng.state625()


def test_full_size_ring_bit_exact_indexes_and_properties():
    """BASELINE config 2: replay 1M, batch 32 — the oracle only needs the terminals array to
    name the indexes, so index parity is checked bit-exactly at full size; the gathered bytes are
    checked through the window property post[:, :3] == pre[:, 1:] and against the tiled base."""
    size, batch, blk = 1_000_000, 32, 10_000
    g = np.random.default_rng(0)
    base = g.integers(0, 256, (blk, 84, 84), dtype=np.uint8)
    actions = g.integers(0, 4, size, dtype=np.uint8)
    rewards = g.integers(-1, 2, size, dtype=np.int64)
    terminals = (g.random(size) < 0.005)
    mem = _mem(size, batch, rng="python")
    for s in range(0, size, blk):
        mem.add_batch(actions[s:s + blk], rewards[s:s + blk], base, terminals[s:s + blk])
    mem.set_cursor(size, 123456)
    orc = ReplayOracle(8, batch_size=batch)                   # tiny shell: borrow the arrays it needs
    orc.size, orc.count, orc.current, orc.terminals = size, size, 123456, terminals
    random.seed(1)
    rng = MT19937.from_python(random)
    for _ in range(25):
        idx = orc.sample_indexes(rng)
        pre, a, r, post, t = mem.getMinibatch()
        assert (mem.last_indexes == idx).all()
        assert (post[:, :3] == pre[:, 1:]).all()
        assert (a == actions[idx]).all() and (r == rewards[idx]).all() and (t == terminals[idx]).all()
        for k in (0, 7, 31):
            assert (post[k] == base[(idx[k] - 3 + np.arange(4)) % blk]).all()
        assert not ((idx >= 123456) & (idx - 4 < 123456)).any()
        assert not np.array([terminals[i - 4:i].any() for i in idx]).any()
    assert list(random.getstate()[1]) == rng.state625()


def test_get_state_and_errors():
    mem = _mem(50, 4)
    with pytest.raises(AssertionError):
        mem.getState(0)                                       # empty ring (replay_memory.py:38)
    for (a, r, s, t) in indexed_episode_stream(3, seed=1):
        mem.add(a, r, s, t)
    with pytest.raises(AssertionError):
        mem.getMinibatch()                                    # count <= history_length (:52)
    with pytest.raises(AssertionError):
        mem.add(0, 0, np.zeros((10, 10), np.uint8), False)    # wrong screen shape (:27)
    assert (decode_frame_tag(mem.getState(1)) == [1, 2, 0, 1]).all()   # wrap-around list path (:46-47)
    mem.add(1, 2.9, np.zeros((84, 84), np.uint8), True)       # float reward truncates like numpy int64
    assert mem.rewards[3] == 2 and mem.terminals[3] and mem.actions[3] == 1


def test_state_buffer_matches_reference_golden_and_oracle():
    from simple_dqn_b200 import StateBuffer
    g = np.load(os.path.join(GOLDEN, "replay_golden.npz"))
    buf, orc = StateBuffer(make_args()), StateBufferOracle()
    assert not np.asarray(buf.getStateMinibatch()).any()
    for i, (_, _, s, _) in enumerate(indexed_episode_stream(10, seed=21)):
        buf.add(s)
        orc.add(s)
        if i in (0, 2, 5):
            assert (buf.getState() == orc.getState()).all()
    assert (decode_frame_tag(buf.getState()) == g["statebuffer/row0_tags"]).all()
    assert crc(np.asarray(buf.getStateMinibatch())) == g["statebuffer/crc"]
    assert buf.getStateMinibatch().shape == (32, 4, 84, 84)
    buf.reset()
    assert not buf.buffer.any()
    with pytest.raises(AssertionError):
        buf.add(np.zeros((3, 3), np.uint8))


def test_stale_device_handle_and_out_of_range_action_are_rejected():
    """a DeviceMinibatch that was overwritten by a later draw must not silently train on the newer
    indexes, and an action >= num_actions in the ring (the reference would raise IndexError at
    deepqnetwork.py:141) must surface as an error instead of an out-of-bounds read."""
    from simple_dqn_b200 import DeepQNetwork, ReplayMemory, Stream
    st = Stream()
    ring = ReplayOracle(600, batch_size=32)
    synthetic_ring(ring, seed=3, block=50, terminal_p=0.01, num_actions=4)
    mem = ReplayMemory(600, make_args(), rng="device", device_minibatch=True, stream=st)
    mem.add_batch(ring.actions, ring.rewards, ring.screens, ring.terminals)
    mem.set_cursor(ring.count, ring.current)
    net = DeepQNetwork(4, make_args(), math_mode="tcgen05", stream=st)
    net.callback = type("CB", (), {"on_train": staticmethod(lambda c: None)})()
    random.seed(3)
    first = mem.getMinibatch()
    second = mem.getMinibatch()
    with pytest.raises(AssertionError):
        net.train(first, 0)                      # overwritten by `second`
    net.train(second, 0)                         # the current handle trains
    bad = ring.actions.copy()
    bad[:] = 9                                   # every action out of range for A = 4
    mem2 = ReplayMemory(600, make_args(), rng="device", device_minibatch=True, stream=st)
    mem2.add_batch(bad, ring.rewards, ring.screens, ring.terminals)
    mem2.set_cursor(ring.count, ring.current)
    with pytest.raises(AssertionError):
        net.train(mem2.getMinibatch(), 0)
