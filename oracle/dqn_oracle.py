"""Nature-DQN train/predict oracle in numpy fp32 (test infrastructure only — see oracle/__init__.py).

PARITY UNPINNED against Neon: the arithmetic of /root/reference/src/deepqnetwork.py
runs inside NervanaSystems/neon (unpinned; shipped snapshots record 1.3.0+344372b),
which is not under /root/reference and cannot be installed here.  This file restates
Neon's published algorithm for the one topology the reference builds, anchored on
the reference's call sites (cited per function), and is cross-checked by
tests/test_oracle_dqn.py against an independent torch-CPU autograd implementation
and against the shipped snapshot weights (KAT in SURVEY §8c).

Conventions (all verified on the shipped snapshots, SURVEY §7 "Layout traps"):
  * conv weights  W[C*R*S, K]  ("CRSK"), cross-correlation, no bias, no padding
  * linear weights W[nout, nin], y = W @ x, no bias
  * activations flatten per sample in (C, H, W) order; input channel c = frame c of
    the 4-frame window (channel 3 = newest frame)
  * states arrive as (N, 4, 84, 84) uint8 and are divided by 255 in fp32
    (deepqnetwork.py:94-100)
"""
import numpy as np

F32 = np.float32

# (R, S, K, stride) for the three conv layers, deepqnetwork.py:83-87
CONV_GEOM = [(8, 8, 32, 4), (4, 4, 64, 2), (3, 3, 64, 1)]
FC_HIDDEN = 512  # deepqnetwork.py:89


def layer_shapes(num_actions, history_length=4, screen=(84, 84)):
    """Neon-layout weight shapes, deepqnetwork.py:77-92 → [(256,32),(512,64),(576,64),(512,3136),(A,512)]."""
    c, (h, w) = history_length, screen
    shapes = []
    for (r, s, k, st) in CONV_GEOM:
        shapes.append((c * r * s, k))
        h, w, c = (h - r) // st + 1, (w - s) // st + 1, k
    shapes.append((FC_HIDDEN, c * h * w))
    shapes.append((num_actions, FC_HIDDEN))
    return shapes


def xavier_init(num_actions, seed, history_length=4, screen=(84, 84)):
    """Neon ``Xavier`` (deepqnetwork.py:79-80): U(-s, s), s = sqrt(3 / fan_in);
    fan_in = W.shape[0] for conv (local=True), W.shape[1] for Affine (local=False).
    Draw order = layer order from one ``numpy.random.RandomState(seed)`` (Neon's be.rng)."""
    rng = np.random.RandomState(seed)
    ws = []
    for i, shp in enumerate(layer_shapes(num_actions, history_length, screen)):
        fan_in = shp[0] if i < 3 else shp[1]
        s = np.sqrt(3.0 / fan_in)
        ws.append(rng.uniform(-s, s, shp).astype(F32))
    return ws


def _im2col(x, r, s, st):
    """x (N,C,H,W) → patches (N*P*Q, C*R*S) in (c, r, s) column order (matches CRSK rows)."""
    n, c, h, w = x.shape
    p, q = (h - r) // st + 1, (w - s) // st + 1
    win = np.lib.stride_tricks.sliding_window_view(x, (r, s), axis=(2, 3))[:, :, ::st, ::st]  # N,C,P,Q,R,S
    cols = np.ascontiguousarray(win.transpose(0, 2, 3, 1, 4, 5)).reshape(n * p * q, c * r * s)
    return cols, p, q


def forward(weights, states_u8, keep=False):
    """deepqnetwork.py:94-100 (_setInput) + Model.fprop for the layers of :77-92.

    Returns q (N, A) float32 (i.e. ``qvalues.T`` of deepqnetwork.py:186); with keep=True
    also the per-layer tensors needed by :func:`backward`.
    """
    x = states_u8.astype(F32) / F32(255.0)                      # be.divide(input, 255, input)
    acts = {"x0": x}
    h = x
    for li, (r, s, k, st) in enumerate(CONV_GEOM):
        cols, p, q = _im2col(h, r, s, st)
        z = cols @ weights[li]                                  # (N*P*Q, K)
        z = np.maximum(z, F32(0))                               # Rectlin
        h = np.ascontiguousarray(z.reshape(h.shape[0], p, q, k).transpose(0, 3, 1, 2))  # N,K,P,Q
        if keep:
            acts["cols%d" % li] = cols
        acts["h%d" % (li + 1)] = h
    flat = h.reshape(h.shape[0], -1)                            # (N, 3136) in (C,H,W) order
    h4 = np.maximum(flat @ weights[3].T, F32(0))                # Affine 512 + Rectlin
    q = h4 @ weights[4].T                                       # Affine A, linear
    acts["flat"], acts["h4"] = flat, h4
    return (q, acts) if keep else q


def td_targets(preq, maxpostq, actions, rewards, terminals, discount=0.99, min_reward=-1, max_reward=1):
    """deepqnetwork.py:133-143: host loop in Python floats (double), stored into a float32 array."""
    targets = preq.copy()
    rewards = np.clip(rewards, min_reward, max_reward)
    for i, a in enumerate(actions):
        if terminals[i]:
            targets[i, a] = float(rewards[i])
        else:
            targets[i, a] = float(rewards[i]) + discount * float(maxpostq[i])
    return targets


def backward(weights, acts, deltas):
    """Model.bprop(deltas) (deepqnetwork.py:162) → list of dW in Neon layout (sum over batch).

    ReLU derivative is (output > 0); the first layer computes no input gradient.
    """
    n = deltas.shape[0]
    grads = [None] * 5
    grads[4] = deltas.T @ acts["h4"]                            # (A, 512)
    d = (deltas @ weights[4]) * (acts["h4"] > 0)                # (N, 512)
    grads[3] = d.T @ acts["flat"]                               # (512, 3136)
    d = d @ weights[3]                                          # (N, 3136)
    for li in (2, 1, 0):
        r, s, k, st = CONV_GEOM[li]
        h_out = acts["h%d" % (li + 1)]                          # N,K,P,Q
        _, _, p, q = h_out.shape
        d = d.reshape(n, k, p, q) * (h_out > 0)
        dz = np.ascontiguousarray(d.transpose(0, 2, 3, 1)).reshape(n * p * q, k)
        grads[li] = acts["cols%d" % li].T @ dz                  # (C*R*S, K)
        if li == 0:
            break
        x_in = acts["h%d" % li]
        c = x_in.shape[1]
        dcols = (dz @ weights[li].T).reshape(n, p, q, c, r, s)
        dx = np.zeros_like(x_in)
        for rr in range(r):
            for ss in range(s):
                dx[:, :, rr:rr + st * p:st, ss:ss + st * q:st] += dcols[:, :, :, :, rr, ss].transpose(0, 3, 1, 2)
        d = dx
    return [g.astype(F32) for g in grads]


def rmsprop_update(weights, states, grads, batch_size, lr=0.00025, decay=0.95, eps=1e-6):
    """Neon ``RMSProp.optimize`` (deepqnetwork.py:51-53, :165), fp32, in place:
    g = dW / bsz;  s = decay*s + (1-decay)*g^2;  W = W - lr*g / (sqrt(s + eps) + eps)."""
    for w, s, g in zip(weights, states, grads):
        g = g / F32(batch_size)
        s[...] = F32(decay) * s + np.square(g) * F32(1.0 - decay)
        w[...] = w - (g * F32(lr)) / (np.sqrt(s + F32(eps)) + F32(eps))


def adam_update(weights, states, grads, batch_size, t, lr=0.00025, beta_1=0.9, beta_2=0.999, eps=1e-8):
    """Neon ``Adam.optimize`` (deepqnetwork.py:54-56, :165) [neon-recall, parity unpinned], fp32, in place;
    ``states[l] = [m, v]``, ``t`` = number of optimize() calls including this one (Adam.t after ``self.t += 1``):
      l = lr * sqrt(1 - beta_2**t) / (1 - beta_1**t)        (Python doubles -> fp32 scalars, fp32 sqrt/divide)
      g = dW / bsz;  m = m*beta_1 + (1-beta_1)*g;  v = v*beta_2 + (1-beta_2)*g*g;  W = W - (l*m) / (sqrt(v) + eps)"""
    l = F32(lr) * np.sqrt(F32(1.0 - beta_2 ** t)) / F32(1.0 - beta_1 ** t)
    for w, (m, v), g in zip(weights, states, grads):
        g = g / F32(batch_size)
        m[...] = m * F32(beta_1) + F32(1.0 - beta_1) * g
        v[...] = v * F32(beta_2) + F32(1.0 - beta_2) * g * g
        w[...] = w - (F32(l) * m) / (np.sqrt(v) + F32(eps))


def adadelta_update(weights, states, grads, batch_size, decay=0.95, eps=1e-6):
    """Neon ``Adadelta.optimize`` (deepqnetwork.py:57-59, :165) [neon-recall, parity unpinned], fp32, in place;
    ``states[l] = [E[g^2], E[dx^2], dx]``:
      g = dW / bsz;  s0 = s0*decay + (1-decay)*g*g;  s2 = sqrt((s1 + eps) / (s0 + eps)) * g;
      s1 = s1*decay + (1-decay)*s2*s2;  W = W - s2"""
    for w, (s0, s1, s2), g in zip(weights, states, grads):
        g = g / F32(batch_size)
        s0[...] = s0 * F32(decay) + F32(1.0 - decay) * g * g
        s2[...] = np.sqrt((s1 + F32(eps)) / (s0 + F32(eps))) * g
        s1[...] = s1 * F32(decay) + F32(1.0 - decay) * s2 * s2
        w[...] = w - s2


OPT_STATES = {"rmsprop": 1, "adam": 2, "adadelta": 3}


class DQNOracle:
    """Drop-in-shaped restatement of ``DeepQNetwork`` (deepqnetwork.py:15-192) on the CPU."""

    def __init__(self, num_actions, batch_size=32, discount_rate=0.99, learning_rate=0.00025,
                 decay_rate=0.95, clip_error=1.0, min_reward=-1, max_reward=1, seed=1, weights=None,
                 states=None, target_steps=10000, optimizer="rmsprop"):
        self.num_actions = num_actions
        self.batch_size = batch_size
        self.discount_rate, self.learning_rate, self.decay_rate = discount_rate, learning_rate, decay_rate
        self.clip_error, self.min_reward, self.max_reward = clip_error, min_reward, max_reward
        self.weights = [w.astype(F32).copy() for w in (weights or xavier_init(num_actions, seed))]
        self.optimizer = optimizer
        if optimizer == "rmsprop":
            self.states = [np.zeros_like(w) if states is None else states[i].astype(F32).copy()
                           for i, w in enumerate(self.weights)]
        else:   # adam: [m, v]; adadelta: [E[g^2], E[dx^2], dx] per layer
            self.states = [[np.zeros_like(w) for _ in range(OPT_STATES[optimizer])] if states is None
                           else [a.astype(F32).copy() for a in states[i]] for i, w in enumerate(self.weights)]
        # deepqnetwork.py:63-73: a separate target model when target_steps != 0 else an alias
        self.target_weights = [w.copy() for w in self.weights] if target_steps else self.weights
        self.train_iterations = 0
        self.callback = None
        self.last = {}

    def update_target_network(self):
        # deepqnetwork.py:102-105
        for t, w in zip(self.target_weights, self.weights):
            t[...] = w

    def predict(self, states_u8):
        # deepqnetwork.py:174-186
        assert states_u8.shape[0] == self.batch_size
        return forward(self.weights, states_u8)

    def train(self, minibatch, epoch=0):
        # deepqnetwork.py:107-172
        prestates, actions, rewards, poststates, terminals = minibatch
        assert prestates.shape == poststates.shape and prestates.ndim == 4
        postq = forward(self.target_weights, poststates)                        # :119-121
        maxpostq = postq.max(axis=1)                                            # :124
        preq, acts = forward(self.weights, prestates, keep=True)                # :128-130
        targets = td_targets(preq, maxpostq, actions, rewards, terminals,
                             self.discount_rate, self.min_reward, self.max_reward)  # :133-146
        deltas = preq - targets                                                 # SumSquared grad (:149)
        cost = F32(np.mean(np.sum(np.square(deltas), axis=1) / F32(2.0)))       # :154, before the clip
        if self.clip_error:
            deltas = np.clip(deltas, -self.clip_error, self.clip_error)         # :158-159
        grads = backward(self.weights, acts, deltas.astype(F32))                # :162
        if self.optimizer == "rmsprop":                                         # :165
            rmsprop_update(self.weights, self.states, grads, prestates.shape[0], self.learning_rate, self.decay_rate)
        elif self.optimizer == "adam":
            adam_update(self.weights, self.states, grads, prestates.shape[0], self.train_iterations + 1,
                        self.learning_rate)
        else:
            adadelta_update(self.weights, self.states, grads, prestates.shape[0], self.decay_rate)
        self.train_iterations += 1                                              # :168
        self.last = dict(preq=preq, postq=postq, targets=targets, deltas=deltas, grads=grads, cost=cost)
        if self.callback:
            self.callback.on_train(cost)                                        # :171-172
        return cost


def load_snapshot(path):
    """Read (W list, RMSProp-state list) from either pickle layout in /root/reference/snapshots
    (SURVEY §5 checkpoint row; old layout documented by src/util/convert_weights.py:10-12)."""
    import pickle
    with open(path, "rb") as f:
        d = pickle.load(f, encoding="latin1")
    if "layer_params_states" in d:
        ls = d["layer_params_states"]
        return [l["params"]["W"] for l in ls], [l["states"][0] for l in ls]
    ls = [l for l in d["model"]["config"]["layers"] if "params" in l]
    return [l["params"]["W"] for l in ls], [l["states"][0] for l in ls]
