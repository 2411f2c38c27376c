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
