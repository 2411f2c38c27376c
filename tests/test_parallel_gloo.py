"""world_size-2 `gloo` test (CPU) of the N>1 sharding rule of DESIGN.md §5: two ranks that each take their
slice of ONE global minibatch, sum gradients with an all-reduce and apply RMSProp with g = sum / (world*B)
must reproduce the single-process oracle trained with batch world*B (indexes bit-exact, weights to fp32
reassociation), and must stay bit-identical to each other."""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import rel_l2
from oracle import dqn_oracle as O
from oracle.mt19937 import MT19937
from oracle.replay_oracle import ReplayOracle, synthetic_ring
from simple_dqn_b200.parallel import broadcast_unique_id, global_batch, rank_slice

B, WORLD, A = 4, 2, 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _ring():
    ring = ReplayOracle(400, batch_size=global_batch(WORLD, B))
    synthetic_ring(ring, seed=2, block=50, terminal_p=0.02)
    return ring


def _worker(rank, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    torch.set_num_threads(1)
    uid = broadcast_unique_id(dist, lambda: bytes(range(128)), rank)
    assert uid == bytes(range(128))
    ring = _ring()                                   # replicated ring
    rng = MT19937.from_python(random.Random(5))      # replicated stream
    net = O.DQNOracle(A, batch_size=B, seed=9)
    for _ in range(3):
        pre, act, rew, post, term = ring.getMinibatch(rng)          # the GLOBAL minibatch, same on every rank
        lo, hi = rank_slice(rank, WORLD, B)
        mb = (pre[lo:hi], act[lo:hi], rew[lo:hi], post[lo:hi], term[lo:hi])
        postq = O.forward(net.target_weights, mb[3])
        preq, acts = O.forward(net.weights, mb[0], keep=True)
        targets = O.td_targets(preq, postq.max(axis=1), mb[1], mb[2], mb[4])
        deltas = np.clip(preq - targets, -1, 1).astype(np.float32)
        grads = O.backward(net.weights, acts, deltas)
        for g in grads:                                              # the one exchange step of the path
            t = torch.from_numpy(g)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        O.rmsprop_update(net.weights, net.states, grads, global_batch(WORLD, B))
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), *net.weights, pos=np.array(rng.state625()[-1]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process_oracle(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    ring = _ring()
    rng = MT19937.from_python(random.Random(5))
    ref = O.DQNOracle(A, batch_size=WORLD * B, seed=9)
    w0 = [w.copy() for w in ref.weights]
    for _ in range(3):
        ref.train(ring.getMinibatch(rng))
    assert int(r0["pos"]) == int(r1["pos"]) == rng.state625()[-1]       # same draw count on every rank
    for l in range(5):
        a, b = r0["arr_%d" % l], r1["arr_%d" % l]
        assert (a == b).all(), "ranks must stay bit-identical"
        assert rel_l2(a - w0[l], ref.weights[l] - w0[l]) <= 1e-3, l


def _gather_worker(rank, port, out_dir):
    """fc1 in the default peer-memory schedule: dW4 is not all-reduced; every rank gathers all ranks'
    rows of its two operands (H3 = `flat`, dZ4) and multiplies over the global minibatch."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    torch.set_num_threads(1)
    ring = _ring()
    rng = MT19937.from_python(random.Random(5))
    net = O.DQNOracle(A, batch_size=B, seed=9)
    pre, act, rew, post, term = ring.getMinibatch(rng)
    lo, hi = rank_slice(rank, WORLD, B)
    postq = O.forward(net.target_weights, post[lo:hi])
    preq, acts = O.forward(net.weights, pre[lo:hi], keep=True)
    targets = O.td_targets(preq, postq.max(axis=1), act[lo:hi], rew[lo:hi], term[lo:hi])
    deltas = np.clip(preq - targets, -1, 1).astype(np.float32)
    reduced = torch.from_numpy(O.backward(net.weights, acts, deltas)[3].copy())
    dist.all_reduce(reduced, op=dist.ReduceOp.SUM)                    # what NCCL / the two-shot exchange deliver
    dz4 = ((deltas @ net.weights[4]) * (acts["h4"] > 0)).astype(np.float32)
    rows_h3 = [torch.zeros(B, acts["flat"].shape[1]) for _ in range(WORLD)]
    rows_dz = [torch.zeros(B, dz4.shape[1]) for _ in range(WORLD)]
    dist.all_gather(rows_h3, torch.from_numpy(np.ascontiguousarray(acts["flat"])))   # k_xpush, channel 0
    dist.all_gather(rows_dz, torch.from_numpy(dz4))                                  # k_xpush, channel 1
    gathered = (torch.cat(rows_dz).T @ torch.cat(rows_h3)).numpy()   # fc1_wgrad over all world x B rows
    np.savez(os.path.join(out_dir, "fc1_rank%d.npz" % rank), reduced=reduced.numpy(), gathered=gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_fc1_gradient_from_gathered_operands_equals_the_reduced_gradient(tmp_path):
    port = _free_port()
    mp.spawn(_gather_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    r0 = np.load(tmp_path / "fc1_rank0.npz")
    r1 = np.load(tmp_path / "fc1_rank1.npz")
    assert (r0["gathered"] == r1["gathered"]).all(), "same operands in the same order: identical bits on every rank"
    assert np.abs(r0["reduced"]).max() > 0
    assert rel_l2(r0["gathered"], r0["reduced"]) <= 1e-6


def test_rank_slice_partition():
    for world in (1, 2, 4, 8):
        spans = [rank_slice(r, world, 32) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == global_batch(world, 32)
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    with pytest.raises(AssertionError):
        rank_slice(2, 2, 32)


def _bcast_worker(rank, port, out_dir):
    from simple_dqn_b200.parallel import ReplicatedReplay
    from simple_dqn_b200.synthetic_env import SyntheticEnvironment
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    ring = ReplayOracle(64, batch_size=4)

    class Ring:      # add_batch on the oracle ring = n consecutive add()s
        def add_batch(self, actions, rewards, screens, terminals):
            for a, r, s, t in zip(actions, rewards, screens, terminals):
                ring.add(int(a), int(r), s, bool(t))
    rep = ReplicatedReplay(Ring(), dist, rank, src=0, block=4)
    env = SyntheticEnvironment(4, seed=1) if rank == 0 else None      # only rank 0 owns an environment
    for t in range(150):                                               # wraps the 64-frame ring; 150 % 4 != 0
        if rank == 0:
            r = env.act(t % 4)
            rep.add(t % 4, r, env.getScreen(), env.isTerminal())
        else:
            rep.add()
    rep.flush()                                                        # the tail that did not fill a block
    np.savez(os.path.join(out_dir, "ring%d.npz" % rank), screens=ring.screens, actions=ring.actions,
             rewards=ring.rewards, terminals=ring.terminals, cursor=np.array([ring.count, ring.current]))
    dist.destroy_process_group()


def test_replica_broadcast_keeps_rings_identical(tmp_path):
    """§8 f3: one environment, W replicas — every rank's ring ends up byte-identical to the single-process ring."""
    from simple_dqn_b200.synthetic_env import SyntheticEnvironment
    port = _free_port()
    mp.spawn(_bcast_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    ref = ReplayOracle(64, batch_size=4)
    env = SyntheticEnvironment(4, seed=1)
    for t in range(150):
        r = env.act(t % 4)
        ref.add(t % 4, r, env.getScreen(), env.isTerminal())
    for rank in range(WORLD):
        g = np.load(os.path.join(str(tmp_path), "ring%d.npz" % rank))
        assert (g["screens"] == ref.screens).all() and (g["actions"] == ref.actions).all()
        assert (g["rewards"] == ref.rewards).all() and (g["terminals"] == ref.terminals).all()
        assert tuple(g["cursor"]) == (ref.count, ref.current)
