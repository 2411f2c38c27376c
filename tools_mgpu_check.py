"""Developer tool: staged 2..N-rank check of the NCCL path with progress prints (run under torchrun)."""
import os, sys, time, random
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench import make_args, synthetic_meta, NUM_ACTIONS
from simple_dqn_b200 import DeepQNetwork, ReplayMemory
from simple_dqn_b200.parallel import broadcast_unique_id
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
def say(*a):
    print("[rank %d %.1fs]" % (rank, time.time() - T0), *a, flush=True)
T0 = time.time()
torch.cuda.set_device(lr)
dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)
say("pg up")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
replay = 20000
base, actions, rewards, terminals = synthetic_meta(replay)
mem = ReplayMemory(replay, make_args(32 * world), device=lr, stream=stream, rng="device")
mem.add_batch(actions, rewards, base[:replay] if len(base) >= replay else np.tile(base, (2, 1, 1))[:replay], terminals)
mem.set_cursor(replay, 1234)
net = DeepQNetwork(NUM_ACTIONS, make_args(32), device=lr, stream=stream, math_mode=os.environ.get("MATH", "tcgen05"))
net.update_target_network()
say("objects up")
uid = broadcast_unique_id(dist, DeepQNetwork.comm_unique_id, rank)
say("uid ok")
net.comm_init(uid, rank, world)
say("comm up")
random.seed(1); mem.seed_device_rng(random)
net.train_fused(mem, 3); torch.cuda.synchronize()
say("3 fused steps ok, costs", net.last_costs(3))
w = net.get_weights(with_states=False)
chk = torch.tensor([float(np.sum([np.abs(x).sum() for x in w]))], dtype=torch.float64)
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
say("weight checksums", [float(c) for c in allc])
assert all(float(c) == float(allc[0]) for c in allc), "ranks diverged"
t = time.time(); net.train_fused(mem, 500); torch.cuda.synchronize(); say("500 steps: %.1f us/step" % ((time.time() - t) / 500 * 1e6))
dist.barrier(); say("done")
