"""Developer tool: GPU timeline of ONE steady-state fused train step inside the replayed CUDA graph (all branches, PDL)."""
import os, sys, random
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_args, synthetic_meta, NUM_ACTIONS
from simple_dqn_b200 import DeepQNetwork, ReplayMemory, Stream, _lib as L
st = Stream()
replay = 50000
base, actions, rewards, terminals = synthetic_meta(replay)
B = int(os.environ.get("BATCH", "32"))
mem = ReplayMemory(replay, make_args(B), stream=st, rng="device")
for s in range(0, replay, 10000):
    mem.add_batch(actions[s:s + 10000], rewards[s:s + 10000], base, terminals[s:s + 10000])
mem.set_cursor(replay, 1234)
net = DeepQNetwork(NUM_ACTIONS, make_args(B), stream=st, math_mode=os.environ.get("MATH", "tcgen05"))
net.update_target_network()
random.seed(1); mem.seed_device_rng(random)
net.train_fused(mem, 50); st.synchronize()
L.ktrace_begin(0, step=12)
net.train_fused(mem, 16); st.synchronize()     # re-captures with timing slots, records the 12th step of the batch
rows = L.ktrace_end()
t0 = min(r[1] for r in rows)
print("%-14s %9s %9s %8s" % ("kernel", "start_us", "end_us", "dur_us"))
for name, a, b in sorted(rows, key=lambda r: r[1]):
    print("%-14s %9.2f %9.2f %8.2f" % (name, (a - t0) / 1e3, (b - t0) / 1e3, (b - a) / 1e3))
print("step span %.2f us" % ((max(r[2] for r in rows) - t0) / 1e3))
