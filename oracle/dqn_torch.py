"""torch-CPU restatement of the train step (test infrastructure / CPU baseline only).

Same algorithm as oracle/dqn_oracle.py (which follows /root/reference/src/deepqnetwork.py:107-172
line by line) but with the conv stack executed by torch's CPU kernels (oneDNN) and the backward
pass by autograd — the strongest plausible CPU implementation of the reference's hot path on the
host cores, used as the `cpu_baseline` / `--impl reference` timing leg of bench.py and as an
independent cross-check of the numpy oracle.  PARITY UNPINNED against Neon (see oracle/__init__.py).
"""
import numpy as np
import torch

from .dqn_oracle import CONV_GEOM, td_targets


class TorchDQN:
    def __init__(self, weights, states=None, discount=0.99, lr=0.00025, decay=0.95, clip_error=1.0,
                 min_reward=-1, max_reward=1):
        self.w = [torch.tensor(np.asarray(w, np.float32)) for w in weights]
        self.s = [torch.zeros_like(w) if states is None else torch.tensor(np.asarray(states[i], np.float32))
                  for i, w in enumerate(self.w)]
        self.tw = [w.clone() for w in self.w]
        self.discount, self.lr, self.decay, self.clip = discount, lr, decay, clip_error
        self.min_reward, self.max_reward = min_reward, max_reward
        self.train_iterations = 0

    @staticmethod
    def _forward(ws, x_u8):
        h = x_u8.float() / 255.0                                              # deepqnetwork.py:94-100
        for li, (r, s, k, st) in enumerate(CONV_GEOM):
            c = h.shape[1]
            w = ws[li].reshape(c, r, s, k).permute(3, 0, 1, 2)               # CRSK -> KCRS
            h = torch.relu(torch.nn.functional.conv2d(h, w, stride=st))
        h = torch.relu(h.flatten(1) @ ws[3].T)
        return h @ ws[4].T

    def update_target_network(self):
        self.tw = [w.clone() for w in self.w]

    def predict(self, states_u8):
        with torch.no_grad():
            return self._forward(self.w, torch.from_numpy(states_u8)).numpy()

    def train(self, minibatch):
        pre, actions, rewards, post, terminals = minibatch
        with torch.no_grad():
            postq = self._forward(self.tw, torch.from_numpy(post))            # :119-121
            maxpostq = postq.max(dim=1).values.numpy()                        # :124
        ws = [w.clone().requires_grad_(True) for w in self.w]
        preq = self._forward(ws, torch.from_numpy(pre))                       # :128-130
        preq_np = preq.detach().numpy()
        targets = td_targets(preq_np, maxpostq, actions, rewards, terminals, self.discount,
                             self.min_reward, self.max_reward)                # :133-146
        deltas = preq_np - targets
        cost = np.float32(np.mean(np.sum(np.square(deltas), axis=1) / 2.0))   # :154
        if self.clip:
            deltas = np.clip(deltas, -self.clip, self.clip)                   # :158-159
        preq.backward(torch.from_numpy(deltas.astype(np.float32)))            # :162
        n = pre.shape[0]
        with torch.no_grad():
            for w, s, wg in zip(self.w, self.s, ws):                          # :165 RMSProp
                g = wg.grad / n
                s.mul_(self.decay).add_(g * g * (1.0 - self.decay))
                w.sub_((g * self.lr) / (torch.sqrt(s + 1e-6) + 1e-6))
        self.train_iterations += 1
        return cost
