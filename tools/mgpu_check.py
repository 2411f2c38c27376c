"""Staged 2..8-rank check of the data-parallel step (peer-memory or NCCL exchange) with progress prints; run under
torchrun (tests/test_gpu_multi.py does).  Env: GRADS=1 (one step, per-layer gradient CRCs, exit), TIMELINE=1
(steady-state in-graph timeline per rank), ORACLE=k (rank 0 also trains the numpy oracle for k steps at the GLOBAL
batch world x 32 on the same ring and MT19937 stream and compares indexes bit for bit and the weight update)."""
import os, sys, time, random, zlib
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_args, synthetic_meta, NUM_ACTIONS
from simple_dqn_b200 import DeepQNetwork, ReplayMemory
from simple_dqn_b200.parallel import broadcast_unique_id
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
def say(*a):
    sys.stdout.write("[rank %d %.1fs] %s\n" % (rank, time.time() - T0, " ".join(str(x) for x in a)))
    sys.stdout.flush()
T0 = time.time()
torch.cuda.set_device(lr)
dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)
say("pg up")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
replay = 20000
base, actions, rewards, terminals = synthetic_meta(replay)
frames = base[:replay] if len(base) >= replay else np.tile(base, (2, 1, 1))[:replay]
mem = ReplayMemory(replay, make_args(32 * world), device=lr, stream=stream, rng="device")
mem.add_batch(actions, rewards, frames, terminals)
mem.set_cursor(replay, 1234)
net = DeepQNetwork(NUM_ACTIONS, make_args(32), device=lr, stream=stream, math_mode=os.environ.get("MATH", "tcgen05"))
ws0 = net.get_weights(with_states=False)
ws0[3] *= np.float32(3); ws0[4] *= np.float32(3)          # Q ~ O(1) like a trained net (as in tests/test_gpu_net.py)
net.set_weights(ws0)
net.update_target_network()
say("objects up")
uid = broadcast_unique_id(dist, DeepQNetwork.comm_unique_id, rank)
say("uid ok")
net.comm_init(uid, rank, world)
say("comm up:", net.comm_status())
random.seed(1); mem.seed_device_rng(random)
if os.environ.get("GRADS"):
    net.train_fused(mem, 1); torch.cuda.synchronize()
    for l, g in enumerate(net.get_grads()):
        torch.cuda.synchronize()
        say("layer %d grad crc %08x  sum %.9g  abs %.9g" % (l, zlib.crc32(g.tobytes()) & 0xffffffff,
                                                             float(g.astype(np.float64).sum()),
                                                             float(np.abs(g.astype(np.float64)).sum())))
    say("status", net.comm_status())
    dist.barrier(); net.comm_destroy(); dist.destroy_process_group(); sys.exit(0)
K = int(os.environ.get("ORACLE", "0"))
if K:
    # ---- N ranks x 32 samples must be ONE step of the single-process reference at batch_size = N * 32
    from oracle import dqn_oracle as O
    from oracle.mt19937 import MT19937
    from oracle.replay_oracle import ReplayOracle
    net.train_fused(mem, K); torch.cuda.synchronize()
    costs = net.last_costs(K)
    idx_dev = np.asarray(mem._download(9, np.int32, (32 * world,)))          # PTR_INDEXES: the last global draw
    w_dev = net.get_weights(with_states=False)
    if rank == 0:
        ring = ReplayOracle(replay, batch_size=32 * world)
        ring.screens[:] = frames; ring.actions[:] = actions; ring.rewards[:] = rewards; ring.terminals[:] = terminals
        ring.count, ring.current = replay, 1234
        random.seed(1)
        rng = MT19937.from_python(random)
        orc = O.DQNOracle(NUM_ACTIONS, batch_size=32 * world, weights=ws0)
        ref_costs = []
        for _ in range(K):
            idx = ring.sample_indexes(rng)
            ref_costs.append(float(orc.train(ring.gather(idx))))
        assert (idx_dev == idx).all(), "global minibatch indexes differ from the single-process draw"
        say("oracle: indexes of the global minibatch bit-exact (%d x %d)" % (world, 32))
        # the product reports each rank's cost over ITS 32 samples; the oracle's is the mean over the global batch
        for l in range(5):
            num = np.linalg.norm(w_dev[l].astype(np.float64) - orc.weights[l])
            den = np.linalg.norm(orc.weights[l].astype(np.float64) - ws0[l])
            say("oracle: layer %d update rel-L2 error %.3e" % (l, num / den))
            # 3 updates from zero RMSProp state are sign-like for small-gradient elements (chaotic: tests/test_gpu_net.py
            # ::test_trajectory_20_steps…): 3e-2 after 3 steps, the single-step bar of 2e-2 is held by the 1-GPU tests
            assert num <= 3e-2 * den, (l, num / den)
        say("oracle: weights after %d steps at global batch %d match (update rel-L2 <= 3e-2); ref costs %s, rank-0 costs %s"
            % (K, 32 * world, ["%.5f" % c for c in ref_costs], ["%.5f" % c for c in costs]))
    allc = [None] * world
    dist.all_gather_object(allc, [float(c) for c in costs])
    if rank == 0:
        mean_costs = np.mean(np.array(allc), axis=0)
        say("oracle: mean of the ranks' costs %s" % ["%.5f" % c for c in mean_costs])
        assert np.allclose(mean_costs, ref_costs, rtol=2e-3), (mean_costs, ref_costs)
else:
    net.train_fused(mem, 3); torch.cuda.synchronize()
    say("3 fused steps ok, costs", net.last_costs(3))
w = net.get_weights(with_states=False)
say("weights crc32 %08x" % (zlib.crc32(b"".join(np.ascontiguousarray(x).tobytes() for x in w)) & 0xffffffff))
chk = torch.tensor([float(np.sum([np.abs(x).sum() for x in w]))], dtype=torch.float64)
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
say("weight checksums", [float(c) for c in allc])
assert all(float(c) == float(allc[0]) for c in allc), "ranks diverged"
t = time.time(); net.train_fused(mem, 500); torch.cuda.synchronize(); say("500 steps: %.1f us/step" % ((time.time() - t) / 500 * 1e6))
# host cost of enqueueing one step (300 graph launches fit the launch queue: the host is not throttled by the device)
for _ in range(3):
    dist.barrier(); torch.cuda.synchronize()
    t = time.time(); net.train_fused(mem, 300); t_enq = time.time() - t; torch.cuda.synchronize(); t_all = time.time() - t
    say("300 steps: host enqueue %.1f us/step, until done %.1f us/step" % (t_enq / 300 * 1e6, t_all / 300 * 1e6))

say("comm status after run:", net.comm_status())
if os.environ.get("TIMELINE"):
    from simple_dqn_b200 import _lib as L
    L.ktrace_begin(lr, step=12)
    net.train_fused(mem, 16); torch.cuda.synchronize()
    rows = L.ktrace_end()
    for rr in range(world):
        dist.barrier()
        if rr == rank and rr < 2:
            t0 = min(r[1] for r in rows if r[1] < 2 ** 63)
            print("---- rank %d steady-state step" % rank)
            for name, a, b in sorted(rows, key=lambda r: r[1]):
                if a < 2 ** 63:
                    print("%-14s %9.2f %9.2f %8.2f" % (name, (a - t0) / 1e3, (b - t0) / 1e3, (b - a) / 1e3))
            sys.stdout.flush()
dist.barrier(); net.comm_destroy(); dist.destroy_process_group(); say("done")
