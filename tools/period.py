"""Developer tool: steady-state period (us per fused train step) of the production graph, device-timed, no tracing."""
import os, sys, random
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_args, synthetic_meta, NUM_ACTIONS
from simple_dqn_b200 import DeepQNetwork, ReplayMemory
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
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
net.train_fused(mem, 300); st.synchronize()
ts = st
res = []
for rep in range(int(os.environ.get("REPS", "5"))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ts)
    net.train_fused(mem, 2000)
    e1.record(ts)
    st.synchronize()
    res.append(e0.elapsed_time(e1) / 2000 * 1e3)
import time
for _ in range(2):
    st.synchronize()
    t = time.time(); net.train_fused(mem, 300); t_enq = time.time() - t; st.synchronize(); t_all = time.time() - t
    print("300 steps: host enqueue %.1f us/step, until done %.1f us/step" % (t_enq / 300 * 1e6, t_all / 300 * 1e6))
print("period_us min %.2f median %.2f  all %s" % (min(res), float(np.median(res)), " ".join("%.2f" % r for r in res)))
