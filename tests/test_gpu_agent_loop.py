"""N2 / BASELINE configs[0], [2]: the reference's control loop driving the PRODUCT classes on the device.

The loop is tests/agent_loop.py::run_restated_loop, proven equal to the reference's converted src/agent.py +
src/statistics.py by tests/test_agent_loop.py.  What is asserted:

  * against the golden trace the REFERENCE loop produced (tests/golden/agent_loop_golden.npz): the position of the
    process-global `random` stream at every phase boundary (i.e. every ε draw, every random action, every
    random-restart length and every index trial of getMinibatch consumed exactly the words the reference
    consumes — §8 a5 end to end), the replay cursor, the ε schedule, every reward / terminal, and every action
    that was drawn at random;
  * against the numpy oracle run in lock-step on the same minibatches (tests/agent_loop.py::LockstepNet): every
    train() cost and every predict() Q row, with the product re-based on the oracle every 4 updates (whole-run
    trace equality between two fp32 implementations is not meaningful for this chaotic system — see LockstepNet);
    greedy actions may differ from the oracle's only inside the measured Q error band;
  * the Statistics pattern (statistics.py:83-90): validation prestates kept by reference and predicted later."""
import numpy as np
import pytest

import agent_loop as AL
from synthetic_env import SyntheticEnvironment
from test_agent_loop import CASES, golden

pytestmark = pytest.mark.gpu


def _product(cfg, num_actions, mode, device_minibatch):
    from simple_dqn_b200 import DeepQNetwork, ReplayMemory, StateBuffer, Stream
    st = Stream()
    mem = ReplayMemory(cfg.replay_size, cfg, rng="python", device_minibatch=device_minibatch, stream=st)
    net = DeepQNetwork(num_actions, cfg, math_mode=mode, stream=st)
    buf = StateBuffer(cfg, stream=st)
    return mem, net, buf


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("mode", ["tcgen05", "fp32"])
def test_reference_loop_on_product_classes(name, mode):
    if mode == "fp32" and name != "pong_repeat2":
        pytest.skip("the SIMT twin runs the short case only")
    spec = CASES[name]
    cfg = AL.loop_config(**spec["cfg"])
    _, _, OracleDQN = AL.oracle_classes()
    env = SyntheticEnvironment(spec["num_actions"], seed=spec["env_seed"])
    mem, net, buf = _product(cfg, env.numActions(), mode, device_minibatch=True)
    checker = OracleDQN(env.numActions(), cfg)
    w0 = net.get_weights(with_states=False)
    assert all((a == b).all() for a, b in zip(w0, checker.weights))          # same Xavier draw (a7)
    ls = AL.LockstepNet(net, checker, resync=4)
    tr = AL.run_restated_loop(env, mem, ls, buf, cfg).arrays()
    ref = golden(name)

    # ---- the reference's decisions that do not depend on fp32 round-off: bit-exact
    assert (tr["rng_crc"] == ref["rng_crc"]).all(), "the host `random` stream left the reference's track"
    assert (tr["mem_cursor"] == ref["mem_cursor"]).all()
    assert np.array_equal(tr["rates"], ref["rates"])
    assert (tr["rewards"] == ref["rewards"]).all() and (tr["terminals"] == ref["terminals"]).all()
    assert tr["costs"].shape == ref["costs"].shape and tr["q_rows"].shape == ref["q_rows"].shape
    # random.random() is consumed identically, so the SAME steps are random / greedy; random ones match exactly
    same = tr["actions"] == ref["actions"]
    assert same[:cfg.random_steps].all()                                      # ε = 1: every action is random
    # ---- numbers: against the lock-step oracle
    ce = np.array(ls.cost_err)
    assert ce[ce[:, 0] == 1, 1].max() <= 1e-3, ce[ce[:, 0] == 1, 1].max()     # first update after a re-base
    assert ce[:, 1].max() <= 3e-2, ce[:, 1].max()                             # up to 4 consecutive product updates
    qe = np.array(ls.q_err)
    assert qe[qe[:, 0] == 0, 1].max() <= 1e-3, qe[qe[:, 0] == 0, 1].max()
    assert qe[:, 1].max() <= 5e-2, qe[:, 1].max()
    for idx, gap, err in ls.ties:                                             # greedy action differs from the oracle's
        assert gap <= 2 * err + 1e-6, (idx, gap, err)
    assert len(ls.ties) <= 0.02 * max(ls.predicts, 1)
    assert net.train_iterations == len(ref["costs"])
    # phase rows: steps / games / rewards are environment facts, weight_updates the schedule
    assert np.allclose(tr["phase_rows"][:, [0, 1, 2, 3, 4, 7]], ref["phase_rows"][:, [0, 1, 2, 3, 4, 7]])


@pytest.mark.parametrize("mode", ["tcgen05"])
def test_host_minibatch_and_device_handle_paths_agree(mode):
    """getMinibatch() -> host arrays -> train() (what the unmodified agent.py sees with device_minibatch=False) and
    the DeviceMinibatch handle path train the same network to the same bits: same kernels, same frames."""
    spec = CASES["pong_repeat2"]
    cfg = AL.loop_config(**dict(spec["cfg"], epochs=1, test_steps=30))
    out = []
    for device_minibatch in (False, True):
        env = SyntheticEnvironment(spec["num_actions"], seed=spec["env_seed"])
        mem, net, buf = _product(cfg, env.numActions(), mode, device_minibatch)
        trace = AL.run_restated_loop(env, mem, net, buf, cfg)
        out.append((trace.arrays(), trace.stats_q, net.get_weights(with_states=False)))
    (ta, sa, wa), (tb, sb, wb) = out
    for k in ("actions", "rng_crc", "mem_cursor"):
        assert (ta[k] == tb[k]).all(), k
    assert (ta["costs"] == tb["costs"]).all()
    # Q rows of the agent's own predicts: identical.  The Statistics predicts run on `validation_states`, which in the
    # reference ALIASES the replay's persistent prestates buffer and silently changes with every later
    # getMinibatch; the host-array path reproduces that, a device handle that nobody looks at does not touch the
    # host buffer — the one documented difference of device_minibatch=True.
    assert sa == sb
    agent_rows = np.setdiff1d(np.arange(len(ta["q_rows"])), np.array(sa, dtype=np.int64))
    assert (ta["q_rows"][agent_rows] == tb["q_rows"][agent_rows]).all()
    assert all((a == b).all() for a, b in zip(wa, wb))


def test_step_host_equals_the_separate_calls():
    """b200dqn_net_step_host (frames of the env steps + train_repeat x (sample, train) in one call) does exactly what
    4 x mem.add + train_repeat x (getMinibatch, train) does: same costs, same weights, same `random` position."""
    import random
    spec = CASES["pong_repeat2"]
    cfg = AL.loop_config(**spec["cfg"])
    rs = np.random.RandomState(5)
    frames = rs.randint(0, 256, (60, 84, 84)).astype(np.uint8)
    acts = rs.randint(0, 6, 60).astype(np.uint8)
    rews = rs.randint(-2, 3, 60).astype(np.int64)
    terms = (rs.rand(60) < 0.05)
    out = []
    for one_call in (False, True):
        mem, net, _ = _product(cfg, 6, "tcgen05", device_minibatch=True)
        costs = []
        net.callback = type("CB", (), {"on_train": staticmethod(costs.append)})()
        mem.add_batch(acts[:40], rews[:40], frames[:40], terms[:40])
        random.seed(99)
        for k in range(5):
            lo = 40 + 4 * k
            random.random()                               # the agent draws between trains (agent.py:50)
            if one_call:
                net.step_host(mem, acts[lo:lo + 4], rews[lo:lo + 4], frames[lo:lo + 4], terms[lo:lo + 4], train_repeat=2)
            else:
                for j in range(lo, lo + 4):
                    mem.add(int(acts[j]), int(rews[j]), frames[j], bool(terms[j]))
                for _ in range(2):
                    net.train(mem.getMinibatch(), 0)
        out.append((np.array(costs, np.float32), net.get_weights(with_states=False), random.getstate(), mem.count, mem.current))
    (ca, wa, ra, na, cura), (cb, wb, rb, nb_, curb) = out
    assert len(ca) == len(cb) == 10 and (ca == cb).all()
    assert all((x == y).all() for x, y in zip(wa, wb))
    assert ra == rb and (na, cura) == (nb_, curb)
