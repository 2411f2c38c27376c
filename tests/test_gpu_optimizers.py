"""§8(f4): the non-default optimizers of src/deepqnetwork.py:54-59 (--optimizer adam | adadelta) on the device.

Same bars as RMSProp in test_gpu_net.py: the update rule itself is BIT-EXACT against the numpy restatement of
Neon's operation order when both are fed the device's own gradient; a whole train step and a short trajectory
agree within the gradient tolerance.  (Neon is absent: the rules are [neon-recall], parity unpinned — see
oracle/dqn_oracle.py::adam_update / adadelta_update.)"""
import numpy as np
import pytest

from helpers import random_minibatch, rel_l2
from oracle import dqn_oracle as O
from test_gpu_net import MODES, SCHEDS, _paired, _stream

pytestmark = pytest.mark.gpu

OPTS = ["adam", "adadelta"]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("sched", SCHEDS)
@pytest.mark.parametrize("opt", OPTS)
def test_update_rule_bit_exact_given_same_gradient(mode, sched, opt):
    net, orc = _paired(4, mode, stream=_stream(sched), optimizer=opt)
    assert net.num_states == O.OPT_STATES[opt]
    for step in range(3):                      # Adam's step scalar l(t) changes every step: check t = 1, 2, 3
        mb = random_minibatch(32, 4, 5 + step)
        w0 = net.get_weights(with_states=False)
        s0 = net.get_states()
        net.train(mb, 0)
        grads = net.get_grads()
        w1 = net.get_weights(with_states=False)
        s1 = net.get_states()
        wr = [w.copy() for w in w0]
        sr = [[a.copy() for a in st] for st in s0]
        if opt == "adam":
            O.adam_update(wr, sr, grads, 32, t=step + 1)
        else:
            O.adadelta_update(wr, sr, grads, 32)
        for l in range(5):
            for k in range(net.num_states):
                assert (s1[l][k] == sr[l][k]).all(), (step, l, k)
            assert (w1[l] == wr[l]).all(), (step, l)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("opt", OPTS)
@pytest.mark.parametrize("batch", [32, 8])
def test_train_step_and_short_trajectory(mode, opt, batch):
    net, orc = _paired(4, mode, batch=batch, stream=_stream("branches"), optimizer=opt)
    w0 = [w.copy() for w in orc.weights]
    for i in range(4):
        mb = random_minibatch(batch, 4, 40 + i)
        if i == 2:
            net.update_target_network()
            orc.update_target_network()
        net.train(mb, 0)
        ref_cost = orc.train(mb)
        cost = float(net.last_costs(1)[0])
        # step 0 is one train step from identical parameters; later steps inherit the (sign-like, hence chaotic)
        # first updates of the two implementations
        assert abs(cost - ref_cost) <= (2e-3 if i == 0 else 2e-2) * abs(ref_cost), (i, cost, ref_cost)
        if i == 0:
            for l, (g, r) in enumerate(zip(net.get_grads(), orc.last["grads"])):
                assert rel_l2(g, r) <= 2e-3, (l, rel_l2(g, r))
    ws = net.get_weights(with_states=False)
    for l in range(5):
        # Adam / Adadelta steps are sign-like for elements with little history (like RMSProp's): L2 of the update
        assert rel_l2(ws[l] - w0[l], orc.weights[l] - w0[l]) <= 5e-2, (l, rel_l2(ws[l] - w0[l], orc.weights[l] - w0[l]))


@pytest.mark.parametrize("opt", OPTS)
def test_checkpoint_keeps_every_state_array(tmp_path, opt):
    net, _ = _paired(4, "fp32", optimizer=opt)
    net.train(random_minibatch(32, 4, 1), 0)
    for layout in ("neon-1.3.0", "pre-1.0"):
        path = str(tmp_path / ("w_%s.prm" % layout))
        net.save_weights(path, layout=layout)
        net2, _ = _paired(4, "fp32", seed=9, optimizer=opt)
        net2.load_weights(path)
        for a, b in zip(net.get_states(), net2.get_states()):
            assert len(a) == len(b) == O.OPT_STATES[opt]
            assert all((x == y).all() for x, y in zip(a, b))
        assert all((x == y).all() for x, y in zip(net.get_weights(with_states=False), net2.get_weights(with_states=False)))
