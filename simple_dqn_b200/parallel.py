"""Host-side helpers for the data-parallel learners of SURVEY §8(e) / DESIGN.md §5.

Every rank holds a full replica of the ring and of the MT19937 stream, draws the same GLOBAL
minibatch of ``world * batch`` indexes (so the draw is bit-identical to the single-process reference
run with that batch size) and trains on its own contiguous slice; the summed gradient is
all-reduced and the identical RMSProp update (``g = sum(dW) / (world * batch)``) is applied on every
rank."""


def rank_slice(rank, world, batch):
    """[start, stop) of rank's samples inside the global minibatch (the C library uses the same rule:
    ``my_idx = d_idx + rank * nb`` in csrc/net.cu::train_on_ring)."""
    assert 0 <= rank < world and batch > 0
    return rank * batch, (rank + 1) * batch


def global_batch(world, batch):
    return world * batch


def broadcast_unique_id(dist, make_id, rank, src=0):
    """Distribute the 128-byte NCCL unique id created on ``src`` with torch.distributed."""
    box = [make_id() if rank == src else None]
    dist.broadcast_object_list(box, src=src)
    assert isinstance(box[0], (bytes, bytearray)) and len(box[0]) == 128
    return bytes(box[0])


class ReplicatedReplay:
    """ReplayMemory.add for data-parallel learners whose replicas must all see the frames of ONE environment
    (SURVEY §8e "add is applied to all replicas"; §8 f3): the rank that owns the environment (`src`) calls
    :meth:`add` exactly like ``ReplayMemory.add`` (/root/reference/src/replay_memory.py:26-34); every `block`
    env steps — and at the latest before anybody samples (:meth:`flush`) — the buffered frames and their
    (action, reward, terminal) travel ONCE through ``dist.broadcast`` and every rank appends them to its own ring
    with one ``add_batch`` (one H2D per rank per block instead of one upload per env step per rank; the other
    ranks never need the environment).  Ranks other than `src` call :meth:`add` with ``None`` arguments (or simply
    :meth:`flush`) at the same points of the loop: the broadcasts are collective.

    Works with anything that has ``add_batch(actions, rewards, screens, terminals)`` — the device ring or, in
    the CPU tests, the oracle ring."""

    def __init__(self, mem, dist, rank, src=0, block=4, dims=(84, 84)):
        import numpy as np
        self.mem, self.dist, self.rank, self.src, self.block = mem, dist, rank, src, int(block)
        self._np = np
        self._frames = np.zeros((self.block,) + tuple(dims), dtype=np.uint8)
        self._meta = np.zeros((self.block, 3), dtype=np.int64)        # action, reward, terminal
        self._n = 0

    def add(self, action=None, reward=None, screen=None, terminal=None):
        if self.rank == self.src:
            assert screen.shape == self._frames.shape[1:]
            self._frames[self._n] = screen
            self._meta[self._n] = (int(action), int(reward), 1 if terminal else 0)
        self._n += 1
        if self._n == self.block:
            self.flush()

    def flush(self):
        """Collective: broadcast what `src` has buffered and append it on every rank."""
        import torch
        n = self._n
        if n == 0:
            return
        frames = torch.from_numpy(self._frames[:n])
        meta = torch.from_numpy(self._meta[:n])
        self.dist.broadcast(frames, src=self.src)
        self.dist.broadcast(meta, src=self.src)
        np = self._np
        self.mem.add_batch(np.ascontiguousarray(self._meta[:n, 0], dtype=np.uint8), self._meta[:n, 1].copy(),
                           self._frames[:n], np.ascontiguousarray(self._meta[:n, 2], dtype=np.uint8))
        self._n = 0
