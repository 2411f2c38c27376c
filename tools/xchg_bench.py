"""Developer tool: microbenchmark of the peer-memory gradient exchange kernel (run under torchrun, N >= 2)."""
import os, sys, time, ctypes as C
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_args, NUM_ACTIONS
from simple_dqn_b200 import DeepQNetwork, _lib as L
from simple_dqn_b200.parallel import broadcast_unique_id
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
net = DeepQNetwork(NUM_ACTIONS, make_args(32), device=lr, stream=stream)
net.comm_init(broadcast_unique_id(dist, DeepQNetwork.comm_unique_id, rank), rank, world)
mode, ok = net.comm_status()
if rank == 0:
    print("comm:", mode, ok, flush=True)
assert mode == "p2p"
NAMES = {0: "weak ld/st", 1: "strong ld", 2: "strong st", 3: "strong ld+st", 4: "no barriers", 8: "barriers only",
         12: "empty"}
def run(l0, l1, flags, blocks, iters=200):
    us, good = C.c_float(), C.c_int()
    dist.barrier()
    L.call("b200dqn_debug_xchg", net._h, l0, l1, flags, blocks, iters, C.byref(us), C.byref(good))
    return us.value, good.value
NAMES[16] = "LL one-shot"
for (l0, l1, what) in ((3, 4, "fc 6.4 MB"), (2, 2, "conv3 147 KB"), (0, 0, "conv1 32 KB"), (0, 2, "conv1-3 311 KB"),
                       (4, 4, "fc2 8 KB")):
    for flags in (0, 12, 16):
        if flags == 16 and l0 != l1:
            continue
        us, good = run(l0, l1, flags, 0)
        if rank == 0:
            print("%-15s %-14s %8.2f us  kat=%d" % (what, NAMES[flags], us, good), flush=True)
print("[rank %d] status" % rank, net.comm_status(), flush=True)
dist.barrier(); net.comm_destroy(); dist.destroy_process_group()
