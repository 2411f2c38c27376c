/*
 * b200dqn.h — C-ABI of libb200dqn.so: the B200-native (sm_100a) replay-and-train hot path
 * behind tambetm/simple_dqn's ReplayMemory / DeepQNetwork / StateBuffer call surface.
 *
 * The reference has no FFI of its own (it is pure Python calling Neon); the boundary is the
 * three Python classes constructed at /root/reference/src/main.py:103-105 and
 * /root/reference/src/agent.py:12.  Each entry point below names the reference interface it
 * replaces (file:line under /root/reference).  A maintainer binds them with ctypes — see
 * INTEGRATION.md for the stub.
 *
 * Conventions
 *   - every function returns 0 on success or a negative B200DQN_E* code; the message for the
 *     calling thread's last failure is b200dqn_last_error().
 *   - plain pointers and sizes only.  `stream` is a cudaStream_t passed as void* (NULL = the
 *     legacy default stream).  Pointers named host_* are host memory, dev_* are device memory.
 *   - calls are asynchronous on `stream` unless the comment says "synchronises".
 *   - objects own their device memory (the 7.06 GB frame ring, weights, activations); getters
 *     expose device pointers for zero-copy interop (torch.as_tensor via __cuda_array_interface__).
 */
#ifndef B200DQN_H
#define B200DQN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200DQN_VERSION 100 /* 0.1.0 */

enum {
  B200DQN_OK = 0,
  B200DQN_EINVAL = -1,   /* bad argument (the reference would raise AssertionError) */
  B200DQN_ECUDA = -2,    /* CUDA runtime/driver failure, text in b200dqn_last_error() */
  B200DQN_ENOTIMPL = -3, /* a reference flag this build does not implement (NotImplementedError) */
  B200DQN_ENCCL = -4,    /* NCCL failure or libnccl.so.2 not loadable */
  B200DQN_ESTATE = -5    /* call sequence error (e.g. sampling an empty ring) */
};

/* math_mode of b200dqn_net_config */
enum {
  B200DQN_MATH_FP32_SIMT = 0, /* CUDA-core fp32 FFMA implicit GEMM: exact-fp32 reference mode          */
  B200DQN_MATH_TCGEN05 = 1    /* tcgen05.mma kind::f16, fp16 hi/lo split operands (3 MMAs), fp32 TMEM */
};

/* optimizer of b200dqn_net_config — src/deepqnetwork.py:50-61 (--optimizer rmsprop|adam|adadelta, main.py:40) */
enum {
  B200DQN_OPT_RMSPROP = 0,  /* RMSProp(learning_rate, decay_rate), epsilon 1e-6; one state array per W       */
  B200DQN_OPT_ADAM = 1,     /* Adam(learning_rate), beta_1 0.9, beta_2 0.999, epsilon 1e-8; states [m, v]    */
  B200DQN_OPT_ADADELTA = 2  /* Adadelta(decay = decay_rate), epsilon 1e-6; states [E[g^2], E[dx^2], dx]      */
};

typedef struct b200dqn_replay b200dqn_replay; /* replaces class ReplayMemory, src/replay_memory.py:6  */
typedef struct b200dqn_net b200dqn_net;       /* replaces class DeepQNetwork, src/deepqnetwork.py:15 */
typedef struct b200dqn_statebuf b200dqn_statebuf; /* replaces class StateBuffer, src/state_buffer.py:3 */

const char* b200dqn_last_error(void);
int b200dqn_version(void);
/* sm count, compute capability and free/total bytes of `device`; fails unless cc == 10.x */
int b200dqn_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* free_bytes,
                        size_t* total_bytes);

/* Plain device->host / host->device copies of library-owned memory (tests, checkpointing);
 * both synchronise `stream`. */
int b200dqn_copy_to_host(int device, void* host_dst, const void* dev_src, size_t bytes, void* stream);
int b200dqn_copy_to_device(int device, void* dev_dst, const void* host_src, size_t bytes, void* stream);

/* A non-default (non-blocking) CUDA stream owned by the library, for callers that do not bring their
 * own (torch.cuda.Stream().cuda_stream works too).  The fused train path only uses CUDA-graph
 * replay and side-stream branches when it is given a non-default stream. */
int b200dqn_stream_create(int device, void** out_stream);
int b200dqn_stream_destroy(int device, void* stream);
int b200dqn_stream_synchronize(int device, void* stream);

/* Per-launch timing with CUDA events (bench.py's roofline leg): between begin and end every kernel
 * the library launches is followed by an event on its stream.  end synchronises the device and
 * returns, in launch order, a 32-byte label and the elapsed ms since the previous event. */
int b200dqn_profile_begin(int device, void* stream);
int b200dqn_profile_end(int max_entries, char* names32, float* ms, int* count);

/* In-graph kernel timeline: arm, run ONE fused step (train_fused re-captures its graph with timing
 * slots), then read [label, first CTA start, last CTA end] (GPU %globaltimer, ns) per launch. */
int b200dqn_ktrace_begin(int device);
/* Same, but record only the step-th fused step after arming (step >= 1; run at least that many): a
 * steady-state step out of a batch, which is what a multi-rank trace needs. */
int b200dqn_ktrace_begin_at(int device, int step);
int b200dqn_ktrace_end(int max_entries, char* names32, unsigned long long* start_ns, unsigned long long* end_ns,
                       int* count);

/* ------------------------------------------------------------------ replay ring --------- */

/* ReplayMemory.__init__(size, args)  — src/replay_memory.py:7-24.
 * Allocates screens[size][h][w] u8, actions u8, rewards i64, terminals u8 and the
 * (batch,hist,h,w) prestates/poststates staging in HBM.  batch is the GLOBAL minibatch. */
int b200dqn_replay_create(int device, int64_t size, int screen_h, int screen_w, int history_length,
                          int batch_size, b200dqn_replay** out);
int b200dqn_replay_destroy(b200dqn_replay* r);

/* ReplayMemory.add(action, reward, screen, terminal) — src/replay_memory.py:26-34.
 * host_screen is (h,w) u8; staged through pinned memory, async H2D.  reward is stored as int64
 * (np.integer), so a float reward must be truncated by the caller exactly as numpy does. */
int b200dqn_replay_add(b200dqn_replay* r, int action, int64_t reward, const uint8_t* host_screen,
                       int terminal, void* stream);
/* n consecutive add() calls in one transfer (ring fill / vectorised actors). */
int b200dqn_replay_add_batch(b200dqn_replay* r, int64_t n, const uint8_t* host_actions,
                             const int64_t* host_rewards, const uint8_t* host_screens,
                             const uint8_t* host_terminals, void* stream);
/* the attributes `count` and `current` (src/replay_memory.py:17-18); set is for tests/benches */
int b200dqn_replay_get_cursor(const b200dqn_replay* r, int64_t* count, int64_t* current);
int b200dqn_replay_set_cursor(b200dqn_replay* r, int64_t count, int64_t current);

/* ReplayMemory.getState(index) — src/replay_memory.py:37-48 (negative / wrap-around indexes
 * included).  Writes (hist,h,w) u8 to host_out; synchronises. */
int b200dqn_replay_get_state(b200dqn_replay* r, int64_t index, uint8_t* host_out, void* stream);

/* The random stream of random.randint at src/replay_memory.py:59: CPython's MT19937 state as
 * returned by random.getstate()[1] (624 key words + position).  set uploads it; get downloads
 * the advanced state (synchronises) so the host can random.setstate() and stay in lock-step. */
int b200dqn_replay_set_rng(b200dqn_replay* r, const uint32_t host_mt625[625], void* stream);
int b200dqn_replay_get_rng(b200dqn_replay* r, uint32_t host_mt625[625], void* stream);
/* set_rng without the stream synchronisation, key words and position passed separately (they are separate fields of
 * CPython's generator object): staged through a small pinned ring, asynchronous. */
int b200dqn_replay_set_rng_parts(b200dqn_replay* r, const uint32_t* host_key624, uint32_t host_pos, void* stream);

/* The sampling loop of ReplayMemory.getMinibatch — src/replay_memory.py:55-69 — on the device:
 * draws py3 randint(hist, count-1) trials from the MT19937 stream, applies the two rejection
 * tests (:61, :65) and keeps the first `batch` accepted indexes in acceptance order.
 * Results stay on the device (B200DQN_PTR_INDEXES, B200DQN_PTR_WORDS_CONSUMED). */
int b200dqn_replay_sample(b200dqn_replay* r, void* stream);
/* Same, and returns how many MT19937 words the draw consumed (waits for the sampler only, by polling a host-mapped
 * word — no memcpy, no stream synchronisation): the host keeps its own
 * `random` in lock-step by discarding that many 32-bit words instead of downloading the 2.5 KB state. */
int b200dqn_replay_sample_sync(b200dqn_replay* r, uint32_t* host_words_consumed, void* stream);
/* Test hook: bypass the sampler and use caller-chosen indexes (host int32[batch]). */
int b200dqn_replay_set_indexes(b200dqn_replay* r, const int32_t* host_indexes, void* stream);

/* The copy half of getMinibatch — src/replay_memory.py:71-78: materialises prestates,
 * poststates, actions, rewards, terminals for the sampled indexes in the device staging
 * buffers (TMA bulk copies, one CTA per sample-frame). */
int b200dqn_replay_gather(b200dqn_replay* r, void* stream);

/* Copy the staged minibatch (any pointer may be NULL) to host arrays shaped as the reference
 * returns them: pre/post (batch,hist,h,w) u8, actions u8[batch], rewards i64[batch],
 * terminals u8[batch] (0/1), indexes i32[batch], words_consumed u32[1].  Synchronises. */
int b200dqn_replay_read_minibatch(b200dqn_replay* r, uint8_t* host_pre, uint8_t* host_actions,
                                  int64_t* host_rewards, uint8_t* host_post, uint8_t* host_terminals,
                                  int32_t* host_indexes, uint32_t* host_words_consumed, void* stream);

enum {
  B200DQN_PTR_SCREENS = 0, B200DQN_PTR_ACTIONS, B200DQN_PTR_REWARDS, B200DQN_PTR_TERMINALS,
  B200DQN_PTR_PRESTATES, B200DQN_PTR_POSTSTATES, B200DQN_PTR_MB_ACTIONS, B200DQN_PTR_MB_REWARDS,
  B200DQN_PTR_MB_TERMINALS, B200DQN_PTR_INDEXES, B200DQN_PTR_WORDS_CONSUMED, B200DQN_PTR_MT_STATE
};
int b200dqn_replay_device_ptr(b200dqn_replay* r, int which, void** dev_ptr, size_t* bytes);

/* ------------------------------------------------------------------ state window -------- */

/* StateBuffer(args) — src/state_buffer.py:9-13: (batch,hist,h,w) u8 zeros on the device. */
int b200dqn_statebuf_create(int device, int screen_h, int screen_w, int history_length, int batch_size,
                            b200dqn_statebuf** out);
int b200dqn_statebuf_destroy(b200dqn_statebuf* s);
/* StateBuffer.add(observation) — src/state_buffer.py:15-18: row 0 shifts left, newest appended. */
int b200dqn_statebuf_add(b200dqn_statebuf* s, const uint8_t* host_screen, void* stream);
/* StateBuffer.reset() — src/state_buffer.py:26-27 */
int b200dqn_statebuf_reset(b200dqn_statebuf* s, void* stream);
/* StateBuffer.getStateMinibatch() / getState() — src/state_buffer.py:20-24; host_out is
 * (batch,hist,h,w) u8 when whole != 0 else (hist,h,w).  Synchronises. */
int b200dqn_statebuf_read(b200dqn_statebuf* s, uint8_t* host_out, int whole, void* stream);
int b200dqn_statebuf_device_ptr(b200dqn_statebuf* s, void** dev_ptr, size_t* bytes);

/* ------------------------------------------------------------------ Q-network ----------- */

typedef struct b200dqn_net_config {
  int num_actions;       /* DeepQNetwork(num_actions, args)        deepqnetwork.py:16-18 */
  int batch_size;        /* args.batch_size (per-rank minibatch)   :19                   */
  int history_length;    /* args.history_length                    :21                   */
  int screen_h, screen_w;/* args.screen_height / width             :22                   */
  double discount_rate;  /* :20  (a Python float: the TD target is formed in double, :141-143) */
  double learning_rate;  /* :51  RMSProp                                                  */
  double decay_rate;     /* :52                                                           */
  double clip_error;     /* :23  (0 disables the clip, as `if self.clip_error:` does)     */
  int min_reward;        /* :24  */
  int max_reward;        /* :25  */
  int target_steps;      /* :65  0 ⇒ the target network aliases the online network (:72-73) */
  int math_mode;         /* B200DQN_MATH_*                                                */
  int optimizer;         /* B200DQN_OPT_*  (:50-61; args.optimizer, main.py:40)           */
} b200dqn_net_config;

int b200dqn_net_config_default(b200dqn_net_config* cfg, int num_actions);

/* DeepQNetwork.__init__ — src/deepqnetwork.py:16-75.  Weights start at zero: the caller
 * initialises them with b200dqn_net_set_weights (Xavier draw or a snapshot). */
int b200dqn_net_create(int device, const b200dqn_net_config* cfg, b200dqn_net** out);
int b200dqn_net_destroy(b200dqn_net* n);

/* Weights cross the boundary in NEON layout: conv W[C*R*S][K], linear W[nout][nin], fp32,
 * C-contiguous (what Model.get_description / the shipped snapshots hold); host_S is the
 * RMSProp state of the same shape (may be NULL).  layer 0..4; which 0 = online, 1 = target.
 * Replaces Model.load_params / save_params — src/deepqnetwork.py:188-192.  Synchronises. */
int b200dqn_net_set_weights(b200dqn_net* n, int which, int layer, const float* host_W, const float* host_S,
                            void* stream);
int b200dqn_net_get_weights(b200dqn_net* n, int which, int layer, float* host_W, float* host_S, void* stream);
int b200dqn_net_layer_shape(const b200dqn_net* n, int layer, int* rows, int* cols);
/* Optimizer state plane k of `layer` (Neon's `states[k]`: RMSProp k = 0; Adam k = 0 m, 1 v; Adadelta k = 0..2),
 * NEON layout like the weights.  b200dqn_net_num_states returns how many planes the configured optimizer keeps.
 * Both synchronise. */
int b200dqn_net_num_states(const b200dqn_net* n, int* count);
int b200dqn_net_set_state(b200dqn_net* n, int which, int layer, int k, const float* host_S, void* stream);
int b200dqn_net_get_state(b200dqn_net* n, int which, int layer, int k, float* host_S, void* stream);

/* DeepQNetwork.update_target_network — src/deepqnetwork.py:102-105 (weights and optimizer state). */
int b200dqn_net_sync_target(b200dqn_net* n, void* stream);

/* DeepQNetwork.predict(states) — src/deepqnetwork.py:174-186.  host_states (batch,hist,h,w) u8,
 * host_q (batch,A) f32 (already transposed as `qvalues.T`).  H2D + forward + D2H; synchronises. */
int b200dqn_net_predict(b200dqn_net* n, const uint8_t* host_states, float* host_q, void* stream);
/* Same on device memory; rows >= live_rows must be all-zero frames (the StateBuffer case,
 * agent.py:55-58): they are not computed — with no biases Q(0) = 0 exactly — and dev_q rows
 * >= live_rows are written as 0.  live_rows = batch computes everything.  Asynchronous. */
int b200dqn_net_predict_device(b200dqn_net* n, const uint8_t* dev_states, int live_rows, float* dev_q,
                               void* stream);

/* The agent's action selection (src/agent.py:55-61) on a device-resident state window: forward for the live rows as ONE
 * captured CUDA graph, Q-values back through host-mapped memory (no memcpy; the call polls).  host_q is (batch, A);
 * rows >= live_rows are exact zeros.  Needs a non-default stream for the graph path (falls back to
 * b200dqn_net_predict_device + copy otherwise).  Synchronises on the result only. */
int b200dqn_net_predict_device_host(b200dqn_net* n, const uint8_t* dev_states, int live_rows, float* host_q,
                                    void* stream);

/* DeepQNetwork.train(minibatch, epoch) — src/deepqnetwork.py:107-172 — from HOST arrays as the
 * reference passes them (drop-in mode).  terminals is u8 0/1.  host_cost receives cost[0,0]
 * (:171); synchronises. */
int b200dqn_net_train(b200dqn_net* n, const uint8_t* host_pre, const uint8_t* host_actions,
                      const int64_t* host_rewards, const uint8_t* host_post, const uint8_t* host_terminals,
                      float* host_cost, void* stream);
/* Same from device-resident minibatch buffers (what b200dqn_replay_gather produced). Async. */
int b200dqn_net_train_device(b200dqn_net* n, const uint8_t* dev_pre, const uint8_t* dev_actions,
                             const int64_t* dev_rewards, const uint8_t* dev_post, const uint8_t* dev_terminals,
                             void* stream);
/* agent.py:112-114 fused: `nsteps` × (getMinibatch sampling → frames read straight from the ring
 * by the first conv layer → train).  No staging copy, no host round trip.  In a multi-GPU
 * communicator every rank samples the same global minibatch and trains on its own slice.
 * Asynchronous; costs land in the device cost ring (b200dqn_net_read_costs). */
int b200dqn_net_train_fused(b200dqn_net* n, b200dqn_replay* r, int nsteps, void* stream);
/* One train step on the indexes ALREADY sampled into the replay object (b200dqn_replay_sample /
 * _set_indexes): the `net.train(mem.getMinibatch())` pair of agent.py:112-114 when getMinibatch
 * returned a device handle.  Frames are read in place from the ring.  Asynchronous. */
int b200dqn_net_train_sampled(b200dqn_net* n, b200dqn_replay* r, void* stream);
/* Same, then returns cost[0,0] of this step in host_cost (written to host-mapped memory by the step's cost kernel; the
 * call polls that word: no memcpy, no stream synchronisation — the rest of the step may still be running) for the
 * `callback.on_train(cost)` of src/deepqnetwork.py:171-172. */
int b200dqn_net_train_sampled_cost(b200dqn_net* n, b200dqn_replay* r, float* host_cost, void* stream);
/* src/agent.py:102-114 as ONE call, for a caller that drives the loop itself (one host->device hop per train step):
 *   nframes x ReplayMemory.add (the env steps since the last train; frames are (h,w) u8, back to back), then
 *   train_repeat x (ReplayMemory.getMinibatch sampling + DeepQNetwork.train) as captured CUDA graphs.
 * host_key624 != NULL: the caller drew from its `random` since the last call — the MT19937 state (624 key words,
 * position) is adopted first.  host_costs (train_repeat floats) / host_words_consumed (MT words the samplings
 * consumed, so the caller can advance its `random`) arrive through host-mapped memory written by the kernels: no
 * memcpy, one wait.  Both NULL: the call is asynchronous.  train_repeat = 0 only appends the frames. */
int b200dqn_net_step_host(b200dqn_net* n, b200dqn_replay* r, int nframes, const uint8_t* host_actions,
                          const int64_t* host_rewards, const uint8_t* host_frames, const uint8_t* host_terminals,
                          int train_repeat, const uint32_t* host_key624, uint32_t host_pos, float* host_costs,
                          uint32_t* host_words_consumed, void* stream);
/* The last `count` (<= 1024) per-step costs, oldest first.  Synchronises. */
int b200dqn_net_read_costs(b200dqn_net* n, int count, float* host_costs, void* stream);
/* train_iterations — src/deepqnetwork.py:168 */
int b200dqn_net_train_iterations(const b200dqn_net* n, int64_t* iters);

enum {
  B200DQN_NET_PTR_Q_ONLINE = 0, /* preq  (batch,A) f32 of the last train/predict — deepqnetwork.py:129 */
  B200DQN_NET_PTR_Q_TARGET,     /* postq (batch,A) f32 — :120                                          */
  B200DQN_NET_PTR_DELTAS,       /* clipped deltas (batch,A) f32 — :159                                 */
  B200DQN_NET_PTR_GRADS,        /* summed dW, internal layout, all layers contiguous                   */
  B200DQN_NET_PTR_WEIGHTS,      /* online fp32 master weights, internal layout                         */
  B200DQN_NET_PTR_COST,         /* device cost ring                                                    */
  B200DQN_NET_PTR_H1,           /* online activations of the last forward, NHWC fp32: (batch,20,20,32) */
  B200DQN_NET_PTR_H2,           /* (batch,9,9,64)                                                      */
  B200DQN_NET_PTR_H3,           /* (batch,7,7,64)                                                      */
  B200DQN_NET_PTR_H4            /* (batch,512)                                                         */
};
int b200dqn_net_device_ptr(b200dqn_net* n, int which, void** dev_ptr, size_t* bytes);
/* The fused optimizers of the tcgen05 engine never materialise dW4; ask them to keep a copy (tests,
 * debugging) before the step whose gradients b200dqn_net_get_grads should return. */
int b200dqn_net_set_keep_grads(b200dqn_net* n, int keep);
/* Last summed gradient of `layer` converted to NEON layout (tests).  Synchronises. */
int b200dqn_net_get_grads(b200dqn_net* n, int layer, float* host_dW, void* stream);
/* Number of kernels one fused train step launches (bench.py's gpu_launches). */
int b200dqn_net_launches_per_step(const b200dqn_net* n, int* launches);

/* Developer aid: clock64() stamps written by CTA (0,0,0) of the tcgen05 kernel whose label equals
 * $B200DQN_TRACE_LABEL (slots documented in csrc/umma2.cuh).  Returns the number of slots or -1. */
int b200dqn_debug_trace(unsigned long long* host_out, int n);

/* ------------------------------------------------------------------ multi-GPU ----------- */

/* Data-parallel learners with replicated replay (new capability, no reference
 * counterpart).  One process per GPU; the 128-byte NCCL unique id is produced on rank 0 and
 * distributed by the host (torch.distributed / a file).  After comm_init every train step
 * all-reduces the summed dW (fp32, 1.69 M elements) over NVLink before the RMSProp update, so
 * weights stay bit-identical on all ranks.  libnccl.so.2 is dlopen()ed at first use (bootstrap,
 * handle exchange, and the fallback data path). */
int b200dqn_comm_unique_id(void* out_id128);
int b200dqn_net_comm_init(b200dqn_net* n, const void* id128, int rank, int world_size);
int b200dqn_net_comm_destroy(b200dqn_net* n);
/* How the gradients travel and whether the exchange is healthy (synchronises the device).
 * *mode: 0 = single learner, 1 = NCCL all-reduce, 2 = peer-memory exchange — every rank maps every
 * other rank's exchange buffers (cudaIpc); fc1's operand rows are gathered with P2P stores so its
 * gradient is never reduced, the small layers use a one-shot LL all-reduce (csrc/comm_p2p.cuh;
 * B200DQN_P2P_SCHED=layer|tail selects the two-shot in-place exchange instead).  Chosen at comm_init
 * when all ranks can map each other, forced off with B200DQN_COMM=nccl.  *error != 0: a peer wait timed out (60 s) since comm_init — the learners
 * are out of step and the results since then are invalid. */
int b200dqn_net_comm_status(b200dqn_net* n, int* mode, int* error);
/* Developer aid (mode 2 only, collective: every rank makes the same call): mean microseconds of `iters`
 * back-to-back in-place exchanges of layers [l0, l1] with the k_xchg switches `flags` and a CTA cap
 * `blocks` (0 = default), and whether one exchange reproduced its known answer.  flags & 16: time the
 * one-shot LL all-reduce of the default schedule instead (l0 == l1, one of layers 0, 1, 2, 4). */
int b200dqn_debug_xchg(b200dqn_net* n, int l0, int l1, int flags, int blocks, int iters, float* us_out, int* ok_out);

#ifdef __cplusplus
}
#endif
#endif /* B200DQN_H */
