"""MT19937 + CPython-3 ``random.randint`` restated (oracle; test infrastructure only).

The reference draws replay indices with ``random.randint(history_length, count - 1)``
(/root/reference/src/replay_memory.py:59).  The generator is CPython's (third-party
to the reference: CPython 3.12 ``Modules/_randommodule.c`` ``genrand_uint32`` and
``Lib/random.py`` ``Random.randrange`` / ``_randbelow_with_getrandbits``):

    randint(a, b)      = a + _randbelow(n),  n = b - a + 1
    _randbelow(n)      : k = n.bit_length(); r = getrandbits(k); while r >= n: r = getrandbits(k)
    getrandbits(k<=32) = genrand_uint32() >> (32 - k)

so every *trial* costs exactly one 32-bit MT19937 output word.  The state handed
around is the 625-tuple of ``random.getstate()[1]``: 624 key words + position.

Pinned by tests/test_oracle_replay.py against CPython's ``random`` itself.
"""
import numpy as np

N = 624
M = 397
MATRIX_A = 0x9908B0DF
UPPER_MASK = 0x80000000
LOWER_MASK = 0x7FFFFFFF


def twist_sequential(mt):
    """In-place regeneration of all 624 words exactly as genrand_uint32 does when pos == N."""
    mt = [int(x) for x in mt]
    for kk in range(N - M):
        y = (mt[kk] & UPPER_MASK) | (mt[kk + 1] & LOWER_MASK)
        mt[kk] = mt[kk + M] ^ (y >> 1) ^ (MATRIX_A if (y & 1) else 0)
    for kk in range(N - M, N - 1):
        y = (mt[kk] & UPPER_MASK) | (mt[kk + 1] & LOWER_MASK)
        mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ (MATRIX_A if (y & 1) else 0)
    y = (mt[N - 1] & UPPER_MASK) | (mt[0] & LOWER_MASK)
    mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ (MATRIX_A if (y & 1) else 0)
    return np.array(mt, dtype=np.uint32)


def twist_segmented(mt):
    """Same regeneration in three dependency-free segments + the last word.

    This is the formulation the CUDA sampler uses: new[i] needs
    old[i], old[i+1] and word (i+397) mod 624 which is *old* for i < 227 and
    *new*[i-227] otherwise, so [0,227), [227,454), [454,623) can each be done
    as one vector step, then word 623 (which needs new[0] and new[396]).
    """
    mt = np.array(mt, dtype=np.uint32).copy()

    def seg(lo, hi):
        i = np.arange(lo, hi)
        y = (mt[i] & np.uint32(UPPER_MASK)) | (mt[i + 1] & np.uint32(LOWER_MASK))
        mag = np.where(y & np.uint32(1), np.uint32(MATRIX_A), np.uint32(0))
        mt[i] = mt[(i + M) % N] ^ (y >> np.uint32(1)) ^ mag

    seg(0, 227)
    seg(227, 454)
    seg(454, 623)
    y = (mt[623] & np.uint32(UPPER_MASK)) | (mt[0] & np.uint32(LOWER_MASK))
    mt[623] = mt[396] ^ (y >> np.uint32(1)) ^ (np.uint32(MATRIX_A) if (int(y) & 1) else np.uint32(0))
    return mt


def temper(y):
    y = int(y)
    y ^= (y >> 11)
    y ^= (y << 7) & 0x9D2C5680
    y ^= (y << 15) & 0xEFC60000
    y ^= (y >> 18)
    return y & 0xFFFFFFFF


class MT19937:
    """Word-level generator carrying CPython's (key[624], pos) state."""

    def __init__(self, state625):
        state625 = [int(x) for x in state625]
        assert len(state625) == N + 1
        self.mt = np.array(state625[:N], dtype=np.uint32)
        self.pos = state625[N]
        self.words_drawn = 0

    @classmethod
    def from_python(cls, rnd):
        """Adopt the state of a ``random.Random`` (or the ``random`` module)."""
        st = rnd.getstate()
        assert st[0] == 3, "CPython MT19937 state version 3 expected"
        return cls(st[1])

    def state625(self):
        return [int(x) for x in self.mt] + [int(self.pos)]

    def to_python(self, rnd):
        """Write the state back into a ``random.Random`` (gauss_next cleared)."""
        rnd.setstate((3, tuple(self.state625()), None))

    def genrand_uint32(self):
        if self.pos >= N:
            self.mt = twist_sequential(self.mt)
            self.pos = 0
        y = self.mt[self.pos]
        self.pos += 1
        self.words_drawn += 1
        return temper(y)

    def randbelow(self, n):
        k = int(n).bit_length()
        assert 0 < k <= 32
        r = self.genrand_uint32() >> (32 - k)
        while r >= n:
            r = self.genrand_uint32() >> (32 - k)
        return r

    def randint(self, a, b):
        return a + self.randbelow(b - a + 1)
