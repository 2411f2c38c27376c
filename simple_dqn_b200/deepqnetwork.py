"""DeepQNetwork with the reference's call surface (/root/reference/src/deepqnetwork.py:15-192):
the Neon model/train/predict replaced by hand-written sm_100a kernels (csrc/net*.cu)."""
import ctypes as C
import logging
import pickle

import numpy as np

from . import _lib as L
from .replay_memory import DeviceMinibatch
from .state_buffer import DeviceStates

logger = logging.getLogger(__name__)

# (R, S, K, stride) of deepqnetwork.py:83-87
_CONV = [(8, 8, 32, 4), (4, 4, 64, 2), (3, 3, 64, 1)]
# --optimizer (main.py:40, deepqnetwork.py:50-61) -> (library code, number of Neon state arrays per W)
_OPTIMIZERS = {"rmsprop": (L.OPT_RMSPROP, 1), "adam": (L.OPT_ADAM, 2), "adadelta": (L.OPT_ADADELTA, 3)}


def _arg(args, name, default):
    return getattr(args, name, default)


class DeepQNetwork:
    def __init__(self, num_actions, args, device=None, math_mode=None, stream=None):
        # remember parameters (:17-26)
        self.num_actions = num_actions
        self.batch_size = args.batch_size
        self.discount_rate = args.discount_rate
        self.history_length = args.history_length
        self.screen_dim = (args.screen_height, args.screen_width)
        self.clip_error = args.clip_error
        self.min_reward = args.min_reward
        self.max_reward = args.max_reward
        self.batch_norm = _arg(args, "batch_norm", False)
        # flags of the reference this build accepts but does not implement (SURVEY §8 a17 note)
        if self.batch_norm:
            raise NotImplementedError("--batch_norm is not implemented on the B200 path")
        self.optimizer = _arg(args, "optimizer", "rmsprop")
        assert self.optimizer in _OPTIMIZERS, "Unknown optimizer"       # :60-61
        if np.dtype(_arg(args, "datatype", "float32")) != np.float32:
            raise NotImplementedError("only --datatype float32 is implemented on the B200 path")
        if _arg(args, "stochastic_round", False):
            raise NotImplementedError("--stochastic_round is not implemented on the B200 path")
        self.device = _arg(args, "device_id", 0) if device is None else device
        self._stream_obj = stream            # keep the stream alive as long as this object uses it
        self._stream = L.stream_ptr(stream)

        cfg = L.NetConfig()
        L.call("b200dqn_net_config_default", C.byref(cfg), num_actions)
        cfg.batch_size = args.batch_size
        cfg.history_length = args.history_length
        cfg.screen_h, cfg.screen_w = self.screen_dim
        cfg.discount_rate = args.discount_rate
        cfg.learning_rate = args.learning_rate
        cfg.decay_rate = args.decay_rate
        cfg.clip_error = float(args.clip_error or 0)
        cfg.min_reward = int(args.min_reward)
        cfg.max_reward = int(args.max_reward)
        cfg.target_steps = int(args.target_steps or 0)
        if math_mode is None:
            math_mode = _arg(args, "math_mode", "fp32")
        cfg.math_mode = {"fp32": L.MATH_FP32_SIMT, "tcgen05": L.MATH_TCGEN05}[math_mode]
        cfg.optimizer, self.num_states = _OPTIMIZERS[self.optimizer]
        self.math_mode = math_mode
        h = C.c_void_p()
        L.call("b200dqn_net_create", self.device, C.byref(cfg), C.byref(h))
        self._h = h

        # model.initialize (:49, :70): Xavier draws from one numpy RandomState(random_seed) —
        # online layers first, then the separately-initialised target model.
        rng = np.random.RandomState(_arg(args, "random_seed", None))
        for which in ((0, 1) if cfg.target_steps else (0,)):
            for layer, shp in enumerate(self.layer_shapes()):
                fan_in = shp[0] if layer < 3 else shp[1]               # Xavier(local=True / False) (:79-80)
                scale = np.sqrt(3.0 / fan_in)
                w = rng.uniform(-scale, scale, shp).astype(np.float32)
                self._set_layer(which, layer, w, np.zeros_like(w))
        self.train_iterations = 0
        self.save_weights_prefix = _arg(args, "save_weights_prefix", None)
        self.callback = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                L.load().b200dqn_net_destroy(h)
            except Exception:
                pass

    # ---- weights in Neon layout
    def layer_shapes(self):
        out = []
        for layer in range(5):
            r, c = C.c_int(), C.c_int()
            L.call("b200dqn_net_layer_shape", self._h, layer, C.byref(r), C.byref(c))
            out.append((r.value, c.value))
        return out

    def _set_layer(self, which, layer, w, s=None):
        w = np.ascontiguousarray(w, dtype=np.float32)
        s = None if s is None else np.ascontiguousarray(s, dtype=np.float32)
        assert w.shape == self.layer_shapes()[layer], (w.shape, self.layer_shapes()[layer])
        L.call("b200dqn_net_set_weights", self._h, which, layer, L.np_ptr(w), L.np_ptr(s), self._stream)

    def set_weights(self, weights, states=None, which=0):
        """states: per layer either ONE array (Neon's ``states[0]``, all RMSProp needs) or the list of the
        optimizer's state arrays (Adam [m, v], Adadelta [E[g^2], E[dx^2], dx])."""
        for layer, w in enumerate(weights):
            st = None if states is None else states[layer]
            if isinstance(st, (list, tuple)):
                self._set_layer(which, layer, w, None)
                self._set_states(which, layer, st)
            else:
                self._set_layer(which, layer, w, st)

    def _set_states(self, which, layer, arrays):
        assert len(arrays) <= self.num_states, (len(arrays), self.num_states)
        for k, a in enumerate(arrays):
            a = np.ascontiguousarray(a, dtype=np.float32)
            assert a.shape == self.layer_shapes()[layer]
            L.call("b200dqn_net_set_state", self._h, which, layer, k, L.np_ptr(a), self._stream)

    def get_weights(self, which=0, with_states=True):
        ws, ss = [], []
        for layer, shp in enumerate(self.layer_shapes()):
            w = np.empty(shp, dtype=np.float32)
            s = np.empty(shp, dtype=np.float32) if with_states else None
            L.call("b200dqn_net_get_weights", self._h, which, layer, L.np_ptr(w), L.np_ptr(s), self._stream)
            ws.append(w)
            ss.append(s)
        return (ws, ss) if with_states else ws

    def get_states(self, which=0):
        """Every optimizer state array per layer, like Neon's ``states`` lists."""
        out = []
        for layer, shp in enumerate(self.layer_shapes()):
            planes = []
            for k in range(self.num_states):
                a = np.empty(shp, dtype=np.float32)
                L.call("b200dqn_net_get_state", self._h, which, layer, k, L.np_ptr(a), self._stream)
                planes.append(a)
            out.append(planes)
        return out

    def keep_grads(self, keep=True):
        """Make the fused optimizers keep a copy of dW so :meth:`get_grads` works (tests / debugging)."""
        L.call("b200dqn_net_set_keep_grads", self._h, int(bool(keep)))

    def get_grads(self):
        out = []
        for layer, shp in enumerate(self.layer_shapes()):
            g = np.empty(shp, dtype=np.float32)
            L.call("b200dqn_net_get_grads", self._h, layer, L.np_ptr(g), self._stream)
            out.append(g)
        return out

    def device_view(self, which, shape):
        p, b = C.c_void_p(), C.c_size_t()
        L.call("b200dqn_net_device_ptr", self._h, which, C.byref(p), C.byref(b))
        return L.DeviceArray(p.value, shape, "<f4", owner=self)

    def _read_f32(self, which, shape):
        return L.download(self.device, self.device_view(which, shape).ptr, shape, np.float32, self._stream)

    def last_q(self):
        """(preq, postq) of the last train() as (batch, A) arrays (deepqnetwork.py:120,129)."""
        shp = (self.batch_size, self.num_actions)
        return self._read_f32(L.NET_PTR_Q_ONLINE, shp), self._read_f32(L.NET_PTR_Q_TARGET, shp)

    def last_activations(self):
        """Online-network activations of the last forward as NCHW arrays like the oracle's."""
        b = self.batch_size
        h1 = self._read_f32(L.NET_PTR_H1, (b, 20, 20, 32)).transpose(0, 3, 1, 2)
        h2 = self._read_f32(L.NET_PTR_H2, (b, 9, 9, 64)).transpose(0, 3, 1, 2)
        h3 = self._read_f32(L.NET_PTR_H3, (b, 7, 7, 64)).transpose(0, 3, 1, 2)
        h4 = self._read_f32(L.NET_PTR_H4, (b, 512))
        return h1, h2, h3, h4

    def last_deltas(self):
        return self._read_f32(L.NET_PTR_DELTAS, (self.batch_size, self.num_actions))

    # ---- reference methods
    def update_target_network(self):
        L.call("b200dqn_net_sync_target", self._h, self._stream)        # :102-105

    def train(self, minibatch, epoch=0):
        """deepqnetwork.py:107-172.  A pristine DeviceMinibatch is trained in place from the ring."""
        if isinstance(minibatch, DeviceMinibatch) and not minibatch.materialised:
            minibatch._check_current()
            mem = minibatch._mem
            cost = C.c_float()
            if not minibatch.sampled:
                # the index draw rides in this step's graph: one launch, results through host-mapped memory
                words = C.c_uint32()
                lockstep = mem.rng_mode == "python"
                key, pos = mem._host_upload_args() if lockstep else (None, 0)
                if not lockstep and not mem._rng_on_device:
                    mem.seed_device_rng()
                want = lockstep or self.callback is not None
                L.call("b200dqn_net_step_host", self._h, mem._h, 0, None, None, None, None, 1, key, pos,
                       C.byref(cost) if want else None, C.byref(words) if lockstep else None, self._stream)
                minibatch.sampled = True
                if lockstep:
                    mem._host_advance(words.value)
                self.train_iterations += 1
                if self.callback:
                    self.callback.on_train(np.float32(cost.value))      # :171-172 (cost[0,0] is a numpy float32)
            elif self.callback:
                L.call("b200dqn_net_train_sampled_cost", self._h, mem._h, C.byref(cost), self._stream)
                self.train_iterations += 1
                self.callback.on_train(np.float32(cost.value))          # :171-172
            else:
                L.call("b200dqn_net_train_sampled", self._h, mem._h, self._stream)
                self.train_iterations += 1
            return
        prestates, actions, rewards, poststates, terminals = minibatch
        assert len(prestates.shape) == 4                                # :110-116
        assert len(poststates.shape) == 4
        assert len(actions.shape) == 1
        assert len(rewards.shape) == 1
        assert len(terminals.shape) == 1
        assert prestates.shape == poststates.shape
        assert prestates.shape[0] == actions.shape[0] == rewards.shape[0] == poststates.shape[0] == terminals.shape[0]
        assert prestates.shape == (self.batch_size, self.history_length) + self.screen_dim
        pre = np.ascontiguousarray(prestates, dtype=np.uint8)
        post = np.ascontiguousarray(poststates, dtype=np.uint8)
        act = np.ascontiguousarray(actions, dtype=np.uint8)
        rew = np.ascontiguousarray(rewards, dtype=np.int64)
        term = np.ascontiguousarray(terminals, dtype=np.uint8)
        cost = C.c_float()
        L.call("b200dqn_net_train", self._h, L.np_ptr(pre), L.np_ptr(act), L.np_ptr(rew), L.np_ptr(post),
               L.np_ptr(term), C.byref(cost), self._stream)
        self.train_iterations += 1                                      # :168
        if self.callback:
            self.callback.on_train(np.float32(cost.value))              # :171-172 (cost[0,0] is a numpy float32)

    def train_fused(self, mem, nsteps=1):
        """`nsteps` x (mem.getMinibatch(); self.train(...)) of agent.py:112-114 with no host round trip
        (device-resident MT19937 stream).  Costs stay on the device — see :meth:`last_costs`."""
        if not mem._rng_on_device:
            mem.seed_device_rng()
        L.call("b200dqn_net_train_fused", self._h, mem._h, int(nsteps), self._stream)
        self.train_iterations += nsteps
        mem._sample_ticket += nsteps
        mem._host_state_in_sync = None        # the device stream ran ahead of the host's `random`

    def step_host(self, mem, actions, rewards, screens, terminals, train_repeat=1):
        """agent.py:102-114 for a caller that owns the loop, as ONE library call: the (action, reward, screen,
        terminal) of the env steps since the last train are appended to `mem`, then `train_repeat` x
        (mem.getMinibatch(); self.train(...)) run on the device.  With ``mem.rng_mode == "python"`` the process-global
        `random` stays in lock-step (state up if it moved, words consumed back).  Returns the costs (float32 array)
        and delivers them to ``callback.on_train`` in order, like the reference's per-train callback."""
        n = len(actions)
        a = np.ascontiguousarray(actions, dtype=np.uint8)
        r = np.ascontiguousarray(rewards, dtype=np.int64)
        s = np.ascontiguousarray(screens, dtype=np.uint8)
        t = np.ascontiguousarray(terminals, dtype=np.uint8)
        assert s.shape == (n,) + tuple(mem.dims) and a.shape == r.shape == t.shape == (n,)
        costs = np.zeros(max(train_repeat, 1), dtype=np.float32)
        words = C.c_uint32()
        lockstep = mem.rng_mode == "python"
        key, pos = mem._host_upload_args() if (lockstep and train_repeat) else (None, 0)
        if not lockstep and not mem._rng_on_device:
            mem.seed_device_rng()
        L.call("b200dqn_net_step_host", self._h, mem._h, n, L.np_ptr(a), L.np_ptr(r), L.np_ptr(s), L.np_ptr(t),
               int(train_repeat), key, pos, L.np_ptr(costs) if train_repeat else None,
               C.byref(words) if (lockstep and train_repeat) else None, self._stream)
        if train_repeat:
            mem._sample_ticket += train_repeat
            if lockstep:
                mem._host_advance(words.value)
            else:
                mem._host_state_in_sync = None
            for c in costs[:train_repeat]:
                self.train_iterations += 1
                if self.callback:
                    self.callback.on_train(np.float32(c))
        return costs[:train_repeat]

    def last_costs(self, count=1):
        out = np.empty(count, dtype=np.float32)
        L.call("b200dqn_net_read_costs", self._h, int(count), L.np_ptr(out), self._stream)
        return out

    def predict(self, states):
        # :176 — the minibatch is full size
        assert tuple(states.shape) == ((self.batch_size, self.history_length,) + self.screen_dim)
        q = np.empty((self.batch_size, self.num_actions), dtype=np.float32)
        if isinstance(states, DeviceStates):
            # one graph launch, Q row(s) back through host-mapped memory (agent.py:55-61 runs this every env step)
            L.call("b200dqn_net_predict_device_host", self._h, C.c_void_p(states.device_ptr()), states.live_rows,
                   L.np_ptr(q), self._stream)
            return q
        st = np.ascontiguousarray(states, dtype=np.uint8)
        L.call("b200dqn_net_predict", self._h, L.np_ptr(st), L.np_ptr(q), self._stream)
        return q                                                        # (batch, A) == qvalues.T (:186)

    def load_weights(self, load_path):
        """Model.load_params (:188-189): both pickle layouts found in the reference's snapshots/ —
        the pre-1.0 ``layer_params_states`` list (breakout/pong; src/util/convert_weights.py:10-12) and the
        neon-1.3.0 ``model.config.layers`` list (seaquest/space_invaders).  Weights AND optimizer states
        are restored (Model.load_params(load_states=True) is Neon's default)."""
        with open(load_path, "rb") as f:
            d = pickle.load(f, encoding="latin1")
        if "layer_params_states" in d:
            ls = d["layer_params_states"]
        else:
            ls = [l for l in d["model"]["config"]["layers"] if "params" in l]
        assert len(ls) == 5, "checkpoint does not hold the five weight layers of deepqnetwork.py:77-92"
        ws = [np.asarray(l["params"]["W"], dtype=np.float32) for l in ls]
        for layer, (l, w) in enumerate(zip(ls, ws)):
            assert w.shape == self.layer_shapes()[layer], \
                "layer %d: checkpoint shape %s, network shape %s" % (layer, w.shape, self.layer_shapes()[layer])
            st = [np.asarray(a, dtype=np.float32) for a in (l.get("states") or [])][:self.num_states]
            self._set_layer(0, layer, w, None)
            if st:
                self._set_states(0, layer, st)
            if len(st) < self.num_states:        # states the checkpoint's optimizer did not keep start at zero
                self._set_states(0, layer, st + [np.zeros_like(w)] * (self.num_states - len(st)))

    def save_weights(self, save_path, layout="neon-1.3.0"):
        """Model.save_params (:191-192).  ``layout="neon-1.3.0"`` (default) writes the structure the reference's
        current Neon writes and reads (same keys, layer list and type strings as snapshots/seaquest_178.pkl);
        ``layout="pre-1.0"`` writes the older ``layer_params_states`` list.  :meth:`load_weights` reads both."""
        ws = self.get_weights(with_states=False)
        ss = self.get_states()
        if layout == "pre-1.0":
            d = {"epoch_index": 0,
                 "layer_params_states": [{"params": {"W": w}, "states": list(s)} for w, s in zip(ws, ss)]}
        else:
            assert layout == "neon-1.3.0", layout
            layers = []
            for i, (w, s) in enumerate(zip(ws, ss)):
                if i < 3:
                    r, _, k, stride = _CONV[i]
                    layers.append({"type": "neon.layers.layer.Convolution",
                                   "config": {"fshape": (r, r, k), "strides": stride, "name": "Convolution_%d" % i,
                                              "parallelism": "Disabled",
                                              "init": {"type": "neon.initializers.initializer.Xavier",
                                                       "config": {"local": True}}},
                                   "params": {"W": w}, "states": list(s)})
                else:
                    layers.append({"type": "neon.layers.layer.Linear",
                                   "config": {"nout": int(w.shape[0]), "name": "Linear_%d" % (i - 3),
                                              "init": {"type": "neon.initializers.initializer.Xavier",
                                                       "config": {"local": False}}},
                                   "params": {"W": w}, "states": list(s)})
                if i < 4:                                            # Rectlin after the first four (:83-89)
                    name = layers[-1]["config"]["name"]
                    layers.append({"type": "neon.layers.layer.Activation",
                                   "config": {"name": name + "_Rectlin",
                                              "transform": {"type": "neon.transforms.activation.Rectlin",
                                                            "config": {"name": "Rectlin_%d" % i}}}})
            d = {"neon_version": "1.3.0+344372b", "epoch_index": 0,
                 "train_input_shape": (self.history_length,) + self.screen_dim,
                 "backend": {"type": "b200dqn", "compat_mode": "neon", "rng_seed": None},
                 "cost": {"type": "neon.layers.layer.GeneralizedCost",
                          "config": {"name": "GeneralizedCost_0",
                                     "costfunc": {"type": "neon.transforms.cost.SumSquared", "config": {}}}},
                 "model": {"type": "neon.layers.container.Sequential", "container": True,
                           "config": {"name": "Sequential_0", "layers": layers}}}
        with open(save_path, "wb") as f:
            pickle.dump(d, f, protocol=2)

    # ---- multi-GPU (new capability)
    def comm_init(self, unique_id, rank, world_size):
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        L.call("b200dqn_net_comm_init", self._h, buf, rank, world_size)

    def comm_destroy(self):
        L.call("b200dqn_net_comm_destroy", self._h)

    def comm_status(self):
        """-> (mode, healthy): mode 'single' | 'nccl' | 'p2p' (peer-memory exchange); synchronises the device.
        healthy is False when a peer wait timed out since comm_init (results since then are invalid)."""
        mode, err = C.c_int(), C.c_int()
        L.call("b200dqn_net_comm_status", self._h, C.byref(mode), C.byref(err))
        return ("single", "nccl", "p2p")[mode.value], err.value == 0

    @staticmethod
    def comm_unique_id():
        buf = (C.c_char * 128)()
        L.call("b200dqn_comm_unique_id", buf)
        return bytes(buf)

    def launches_per_step(self):
        n = C.c_int()
        L.call("b200dqn_net_launches_per_step", self._h, C.byref(n))
        return n.value
