"""A deterministic stand-in for ALEEnvironment (§8 f3: "a synthetic / ALE-compatible Environment"; ALE and OpenCV
are not installed, BASELINE configs[0]/[2]/[4] name Atari games).  Implements the Environment interface of
/root/reference/src/environment.py:7-33 — numActions, restart, act, getScreen, isTerminal, setMode — and nothing
else, so the reference's Agent drives it unchanged (tests/test_agent_loop.py).

Design choices that matter for a parity test:
  * it never touches Python's `random` (ALE has its own generator), so the process-global stream is consumed only
    by the agent and the replay sampler, exactly as with the real emulator;
  * screens, rewards and episode ends are functions of the step counter alone, NOT of the action taken: a legitimate
    last-ulp difference between two fp32 implementations can then flip at most the one stored action of a near-tie,
    it cannot cascade into a different game;
  * rewards include values outside [-1, 1] (the reward clip of deepqnetwork.py:136 must act) and episodes end every
    ~120 steps (terminal handling of replay_memory.py:65 and agent.py:77-79)."""
import numpy as np


class SyntheticEnvironment:
    def __init__(self, num_actions=4, seed=0, dims=(84, 84), episode_mean=120):
        self.num_actions = int(num_actions)
        self.seed = int(seed)
        self.dims = tuple(dims)
        self.episode_mean = int(episode_mean)
        self.t = 0                    # frames emitted so far
        self.mode = "train"
        self._terminal = False
        self._screen = self._frame(0)

    def _h(self, t, salt):
        m = (1 << 64) - 1
        x = (t * 0x9E3779B97F4A7C15 + self.seed * 1000003 + salt) & m      # splitmix64-style scramble of the counter
        x ^= x >> 31
        x = (x * 0xBF58476D1CE4E5B9) & m
        x ^= x >> 29
        return x

    def _frame(self, t):
        g = np.random.Generator(np.random.PCG64([self.seed, t]))
        f = g.integers(0, 256, self.dims, dtype=np.uint8)
        f[(t * 7) % self.dims[0], :] = 255          # a moving bright row: consecutive frames are correlated
        return f

    # ---- environment.py:10-33
    def numActions(self):
        return self.num_actions

    def restart(self):
        self._terminal = False
        self.t += 1
        self._screen = self._frame(self.t)

    def act(self, action):
        assert 0 <= int(action) < self.num_actions
        self.t += 1
        self._screen = self._frame(self.t)
        h = self._h(self.t, 1)
        self._terminal = (h % self.episode_mean) == 0
        return (-2, -1, 0, 0, 0, 0, 1, 3)[(h >> 20) % 8]

    def getScreen(self):
        return self._screen

    def isTerminal(self):
        return self._terminal

    def setMode(self, mode):
        self.mode = mode
