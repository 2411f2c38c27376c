"""Developer tool: a short run of fused train steps for ncu (launch list / --set full captures)."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_args, synthetic_meta, NUM_ACTIONS
from simple_dqn_b200 import DeepQNetwork, ReplayMemory, Stream
st = Stream()
B = int(os.environ.get("BATCH", "32"))
replay = 200000          # 1.4 GB of frames: sampled windows come from HBM, not from L2
base, actions, rewards, terminals = synthetic_meta(replay)
mem = ReplayMemory(replay, make_args(B), stream=st, rng="device")
for s in range(0, replay, 10000):
    mem.add_batch(actions[s:s + 10000], rewards[s:s + 10000], base, terminals[s:s + 10000])
mem.set_cursor(replay, 1234)
net = DeepQNetwork(NUM_ACTIONS, make_args(B), stream=st, math_mode="tcgen05")
net.update_target_network()
random.seed(1); mem.seed_device_rng(random)
net.train_fused(mem, int(os.environ.get("STEPS", "40"))); st.synchronize()
print("costs", net.last_costs(3))
