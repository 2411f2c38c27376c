"""CPU cross-check of the optimizer restatements in oracle/dqn_oracle.py (the rules csrc/optim.cuh implements, which the
GPU tests compare bit for bit against this file) with torch.optim's independent implementations of the same published
algorithms.  Neon (the reference's dependency, deepqnetwork.py:50-61) is absent, so the rules are [neon-recall]; torch
differs from Neon only in where epsilon sits, which bounds the agreement stated in each test."""
import numpy as np
import pytest
import torch

from oracle import dqn_oracle as O

F32 = np.float32


def _problem(seed, scale):
    rng = np.random.RandomState(seed)
    w = [rng.uniform(-1, 1, size=s).astype(F32) for s in ((7, 5), (11,))]
    grads = [[(scale * rng.normal(size=x.shape)).astype(F32) for x in w] for _ in range(6)]
    return w, grads


def _run_torch(opt_cls, w0, grads, bsz, **kw):
    params = [torch.nn.Parameter(torch.tensor(x.copy())) for x in w0]
    opt = opt_cls(params, **kw)
    for step in grads:
        for p, g in zip(params, step):
            p.grad = torch.tensor(g / F32(bsz))
        opt.step()
    return [p.detach().numpy() for p in params]


def _rel(a, b):
    return max(float(np.max(np.abs(x - y)) / (np.max(np.abs(y)) + 1e-30)) for x, y in zip(a, b))


def test_adam_is_kingma_ba():
    """Neon: l = lr*sqrt(1-b2^t)/(1-b1^t); W -= l*m/(sqrt(v)+eps).  torch: the bias-corrected form.  Identical up to
    eps -> eps*sqrt(1-b2^t) in the denominator, i.e. to ~eps/|g| relative."""
    w0, grads = _problem(1, 3.0)
    w = [x.copy() for x in w0]
    st = [[np.zeros_like(x), np.zeros_like(x)] for x in w]
    for t, g in enumerate(grads, 1):
        O.adam_update(w, st, g, 32, t, lr=0.01)
    ref = _run_torch(torch.optim.Adam, w0, grads, 32, lr=0.01, betas=(0.9, 0.999), eps=1e-8)
    assert _rel(w, ref) < 1e-5


def test_adadelta_is_zeiler():
    """Same rule in both libraries (rho = decay, lr = 1): only fp32 rounding order differs."""
    w0, grads = _problem(2, 3.0)
    w = [x.copy() for x in w0]
    st = [[np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)] for x in w]
    for g in grads:
        O.adadelta_update(w, st, g, 32, decay=0.95, eps=1e-6)
    ref = _run_torch(torch.optim.Adadelta, w0, grads, 32, lr=1.0, rho=0.95, eps=1e-6)
    assert _rel(w, ref) < 2e-6


def test_rmsprop_is_tieleman_hinton_up_to_epsilon():
    """Neon: W -= lr*g/(sqrt(s+eps)+eps); torch: lr*g/(sqrt(s)+eps).  The inner epsilon (1e-6) matters wherever
    s <~ 1e-6, so the comparison uses gradients bounded away from zero (|g|/bsz in [0.06, 0.125] => s >= 1.8e-4): there
    the two steps differ by ~eps/(2 s) < 0.3 %."""
    w0, grads = _problem(3, 3.0)
    rng = np.random.RandomState(33)
    grads = [[(rng.uniform(2, 4, size=x.shape) * rng.choice([-1.0, 1.0], size=x.shape)).astype(F32) for x in w0]
             for _ in range(6)]
    w = [x.copy() for x in w0]
    st = [np.zeros_like(x) for x in w]
    for g in grads:
        O.rmsprop_update(w, st, g, 32, lr=0.00025, decay=0.95, eps=1e-6)
    ref = _run_torch(torch.optim.RMSprop, w0, grads, 32, lr=0.00025, alpha=0.95, eps=1e-6)
    moved = [x - y for x, y in zip(w, w0)]
    moved_ref = [x - y for x, y in zip(ref, w0)]
    assert _rel(moved, moved_ref) < 5e-3
    assert _rel(moved, moved_ref) > 1e-5   # ... and they do differ: the inner epsilon is Neon's, not torch's


@pytest.mark.parametrize("name", ["rmsprop", "adam", "adadelta"])
def test_state_plane_counts(name):
    """OPT_STATES is what b200dqn_net_num_states reports and what the checkpoint writer emits per layer."""
    assert O.OPT_STATES[name] == {"rmsprop": 1, "adam": 2, "adadelta": 3}[name]
