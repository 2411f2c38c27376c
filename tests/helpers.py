"""Shared test helpers (CPU + GPU tests)."""
import types
import zlib

import numpy as np


def make_args(**kw):
    """The argparse namespace of /root/reference/src/main.py:16-84 restricted to the fields the
    three hot-path classes read, with the reference's defaults."""
    d = dict(screen_height=84, screen_width=84, history_length=4, batch_size=32, discount_rate=0.99,
             learning_rate=0.00025, decay_rate=0.95, clip_error=1, min_reward=-1, max_reward=1, batch_norm=False,
             backend="gpu", random_seed=7, device_id=0, datatype="float32", stochastic_round=False,
             optimizer="rmsprop", target_steps=10000, save_weights_prefix=None)
    d.update(kw)
    return types.SimpleNamespace(**d)


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) /
                 max(np.linalg.norm(np.asarray(b, np.float64)), 1e-30))


def random_minibatch(n, num_actions, seed, terminal_p=0.3, reward_range=(-3, 4)):
    rs = np.random.RandomState(seed)
    pre = rs.randint(0, 256, (n, 4, 84, 84)).astype(np.uint8)
    post = rs.randint(0, 256, (n, 4, 84, 84)).astype(np.uint8)
    return (pre, rs.randint(0, num_actions, n).astype(np.uint8),
            rs.randint(reward_range[0], reward_range[1], n).astype(np.int64), post, rs.rand(n) < terminal_p)
