"""Replay-ring + state-window oracle (numpy; test infrastructure only — see oracle/__init__.py).

Restates, function by function, /root/reference/src/replay_memory.py and
/root/reference/src/state_buffer.py.  PINNED: tests/test_oracle_replay.py runs the
unmodified reference files side by side with this restatement (when
/root/reference is present) and tests/golden/replay_golden.npz holds outputs the
reference itself produced (generator: tests/golden/make_golden.py).

Differences from the reference are only in *how the random stream is named*: the
reference pulls from the process-global ``random`` module; the oracle pulls from an
explicit :class:`oracle.mt19937.MT19937`, which is word-for-word the same stream
(see that module), and additionally reports how many 32-bit words were consumed so a
device sampler can hand the stream position back to the host.
"""
import numpy as np

from .mt19937 import MT19937


class ReplayOracle:
    def __init__(self, size, screen_height=84, screen_width=84, history_length=4, batch_size=32):
        # replay_memory.py:7-24
        self.size = int(size)
        self.actions = np.zeros(self.size, dtype=np.uint8)
        self.rewards = np.zeros(self.size, dtype=np.int64)        # np.integer == int64 (:11)
        self.screens = np.zeros((self.size, screen_height, screen_width), dtype=np.uint8)
        self.terminals = np.zeros(self.size, dtype=np.bool_)
        self.history_length = history_length
        self.dims = (screen_height, screen_width)
        self.batch_size = batch_size
        self.count = 0
        self.current = 0
        self.prestates = np.zeros((batch_size, history_length) + self.dims, dtype=np.uint8)
        self.poststates = np.zeros((batch_size, history_length) + self.dims, dtype=np.uint8)

    def add(self, action, reward, screen, terminal):
        # replay_memory.py:26-34
        assert screen.shape == self.dims
        self.actions[self.current] = action
        self.rewards[self.current] = reward          # float rewards truncate toward zero on store
        self.screens[self.current, ...] = screen
        self.terminals[self.current] = terminal
        self.count = max(self.count, self.current + 1)
        self.current = (self.current + 1) % self.size

    def getState(self, index):
        # replay_memory.py:37-48
        assert self.count > 0
        index = index % self.count
        h = self.history_length
        if index >= h - 1:
            return self.screens[(index - (h - 1)):(index + 1), ...]
        indexes = [(index - i) % self.count for i in reversed(range(h))]
        return self.screens[indexes, ...]

    def accept(self, index):
        """The two rejection tests of replay_memory.py:61 and :65."""
        h = self.history_length
        if index >= self.current and index - h < self.current:
            return False
        if self.terminals[(index - h):index].any():
            return False
        return True

    def sample_indexes(self, rng: MT19937, batch_size=None):
        """replay_memory.py:55-73 without the copies: accepted indexes in acceptance order."""
        assert self.count > self.history_length
        bs = self.batch_size if batch_size is None else batch_size
        indexes = []
        while len(indexes) < bs:
            index = rng.randint(self.history_length, self.count - 1)
            if self.accept(index):
                indexes.append(index)
        return np.array(indexes, dtype=np.int64)

    def gather(self, indexes):
        """replay_memory.py:71-78 for a given list of accepted indexes."""
        for k, index in enumerate(indexes):
            self.prestates[k, ...] = self.getState(index - 1)
            self.poststates[k, ...] = self.getState(index)
        actions = self.actions[indexes]
        rewards = self.rewards[indexes]
        terminals = self.terminals[indexes]
        return self.prestates, actions, rewards, self.poststates, terminals

    def getMinibatch(self, rng: MT19937):
        # replay_memory.py:50-79
        return self.gather(self.sample_indexes(rng))


class StateBufferOracle:
    """state_buffer.py:3-27."""

    def __init__(self, screen_height=84, screen_width=84, history_length=4, batch_size=32):
        self.history_length = history_length
        self.dims = (screen_height, screen_width)
        self.batch_size = batch_size
        self.buffer = np.zeros((batch_size, history_length) + self.dims, dtype=np.uint8)

    def add(self, observation):
        assert observation.shape == self.dims
        self.buffer[0, :-1] = self.buffer[0, 1:]
        self.buffer[0, -1] = observation

    def getState(self):
        return self.buffer[0]

    def getStateMinibatch(self):
        return self.buffer

    def reset(self):
        self.buffer *= 0


def synthetic_ring(oracle: ReplayOracle, seed=0, block=None, terminal_p=0.005, num_actions=4,
                   count=None, current=None):
    """Fill a ring with the synthetic content of SURVEY §8(d) / BASELINE.md §2.

    A ``block``-frame uniform-random block is tiled through the ring (block=None →
    every frame independent).  Returns nothing; mutates ``oracle``.
    """
    g = np.random.default_rng(seed)
    n = oracle.size
    h, w = oracle.dims
    blk = n if block is None else min(block, n)
    base = g.integers(0, 256, (blk, h, w), dtype=np.uint8)
    for s in range(0, n, blk):
        e = min(n, s + blk)
        oracle.screens[s:e] = base[: e - s]
    oracle.actions[:] = g.integers(0, num_actions, n, dtype=np.uint8)
    oracle.rewards[:] = g.integers(-1, 2, n, dtype=np.int64)
    oracle.terminals[:] = g.random(n) < terminal_p
    oracle.count = n if count is None else count
    oracle.current = (n // 8 + 7) % n if current is None else current


def indexed_episode_stream(n_steps, dims=(84, 84), seed=0, terminal_p=0.02, num_actions=4, block=64):
    """Deterministic (action, reward, screen, terminal) stream used by the golden fixtures.

    Frame t is a random base frame (``block`` distinct bases, tiled) whose first four
    bytes are overwritten with t little-endian, so a gathered state names the ring slot
    it came from — that is how tests recover the *indexes* the reference drew, which
    ``getMinibatch`` itself never returns (replay_memory.py:79).
    """
    g = np.random.default_rng(seed)
    base = g.integers(0, 256, (block,) + tuple(dims), dtype=np.uint8)
    actions = g.integers(0, num_actions, n_steps, dtype=np.uint8)
    rewards = g.integers(-3, 4, n_steps, dtype=np.int64)
    terminals = g.random(n_steps) < terminal_p
    for t in range(n_steps):
        screen = base[t % block].copy()
        screen.reshape(-1)[:4] = np.frombuffer(np.uint32(t).tobytes(), dtype=np.uint8)
        yield int(actions[t]), int(rewards[t]), screen, bool(terminals[t])


def decode_frame_tag(frames):
    """Inverse of the tag written by :func:`indexed_episode_stream`; frames (..., H, W) uint8."""
    flat = np.ascontiguousarray(frames.reshape(frames.shape[:-2] + (-1,))[..., :4])
    return flat.view(np.uint32)[..., 0].astype(np.int64)
