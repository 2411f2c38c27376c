// net.cuh — the Q-network object: fp32 master weights (online + target), RMSProp state,
// activations, gradient partials, all resident in HBM.
#pragma once
#include "common.cuh"
#include "net_simt.cuh"
#include "replay.cuh"
#include "comm_p2p.cuh"
#include "optim.cuh"

namespace b200 {

constexpr int kLayers = 5;
constexpr int kMaxActions = 32;
constexpr int kCostRing = 1024;
constexpr int kHostCosts = 60;            // per-step costs mirrored in host-mapped memory
constexpr int kHostQ = 64, kHostQFloats = 960;   // word offset / capacity of the Q rows in the host-mapped block
constexpr int kFc1Splits = 14;            // 3136 / 14 = 224 = 14 * 16
constexpr int kFc1Chunk = kFlat / kFc1Splits;

struct LayerTable {
  int64_t off[kLayers + 1];   // element offsets into the all-layer parameter vector
  int rows[kLayers], cols[kLayers];  // internal [K][N] shape
  int64_t part_off[kLayers];  // element offsets into the split-K partial scratch
  int splits[kLayers];
};

}  // namespace b200

struct b200dqn_net {
  int device = 0;
  int sm_count = 148;      // queried at create; sizes the capped elementwise grids
  b200dqn_net_config cfg{};
  int nb = 0;  // per-rank minibatch
  int A = 0;
  b200::LayerTable lt{};
  int64_t n_params = 0;

  // parameters (internal layout, all layers contiguous)
  float* d_w = nullptr;   // online weights
  float* d_s = nullptr;   // online optimizer state: n_states planes of n_params (RMSProp 1, Adam 2, Adadelta 3)
  int n_states = 1;
  float* d_optscal = nullptr;  // [0] Adam's step scalar l of the current step (written by the head kernel);
                               // [1] (as u32) fast-path predicts completed
  uint32_t* d_optscal_u32() const { return reinterpret_cast<uint32_t*>(d_optscal) + 1; }
  float* d_tw = nullptr;  // target weights (== d_w when target_steps == 0)
  float* d_ts = nullptr;  // target optimizer state (copied for fidelity with :102-105)
  float* d_g = nullptr;   // summed gradients (all-reduce buffer / get_grads)
  float* d_part = nullptr;  // split-K partials
  int64_t part_elems = 0;

  // activations: [0] online, [1] target
  float* d_h1[2] = {}, *d_h2[2] = {}, *d_h3[2] = {}, *d_h4[2] = {};
  float* d_fc1part = nullptr;  // [2*splits][nb][512]
  float* d_q[2] = {};          // [nb][A]
  float* d_delta = nullptr;    // [nb][A] clipped
  float* d_dz4 = nullptr, *d_dz3 = nullptr, *d_dz2 = nullptr, *d_dz1 = nullptr;
  float* d_cost = nullptr;     // cost ring [kCostRing]
  uint32_t* d_step = nullptr;  // device step counter (cost ring cursor)
  float* d_rowcost = nullptr;    // [nb] per-sample cost
  // host-mapped result block (4 KB): [0] train steps completed (published last, after a system fence), [1]
  // action-range flag, [2] predicts completed, [4 + (step % kHostCosts)] cost of that step (k_cost_finish);
  // [64 ..) Q rows of the last fast-path predict (k_publish_q)
  volatile uint32_t* h_res = nullptr;
  uint32_t predicts_launched = 0;
  cudaGraphExec_t graph_predict_exec = nullptr;   // forward on a device-resident state window + result publish
  const uint8_t* graph_predict_states = nullptr;
  int graph_predict_rows = 0;
  cudaStream_t graph_predict_stream = nullptr;
  b200dqn_replay* step_replay = nullptr;   // set while a step that samples from a ring is being enqueued / captured

  // unfused-mode staging (host minibatch -> device)
  uint8_t* d_pre = nullptr, *d_post = nullptr, *d_act = nullptr, *d_term = nullptr;
  int64_t* d_rew = nullptr;
  int32_t* d_iota1 = nullptr;  // b
  int32_t* d_iota4 = nullptr;  // hist * b
  uint8_t* h_pin = nullptr;    // pinned staging for predict/train host entry points
  size_t pin_bytes = 0;

  int64_t train_iterations = 0;

  // step scheduling: side streams / events for the independent wgrad + optimizer branches, and the
  // captured CUDA graph of one fused step
  cudaStream_t side[4] = {};   // three wgrad/optimizer branches + the collective stream
  cudaEvent_t ev[17] = {};
  bool use_graph = true, use_branches = true;
  bool keep_grads = false;   // fused optimizers also write dW for b200dqn_net_get_grads (tests)
  cudaGraphExec_t graph_exec = nullptr;
  b200dqn_replay* graph_replay = nullptr;
  cudaStream_t graph_stream = nullptr;
  int graph_world = 0;
  int graph_trace_gen = 0;
  int graph_launches = 0;   // kernels launched by one captured step
  // Software-pipelined fc1 update (multi-step train_fused, opt-in B200DQN_DEFER_FC1=1 — measured slower, see
  // b200dqn_net_train_fused): the graph of step t applies step t-1's fc1 update on a
  // side branch under its own forward convolutions (joined before fc1_fwd) instead of under the dgrad chain, where its
  // 296 CTAs compete with the tcgen05 kernels for registers and SM slots; the last step's update is flushed by
  // train_fused before it returns, so nothing outside that call ever sees a pending update.
  cudaGraphExec_t graph_def_exec = nullptr;   // the deferred-update variant of the step graph (same cache keys)
  int graph_def_launches = 0;
  bool defer_fc1 = false;                     // set while that variant is being captured
  bool fc1_pending = false;                   // host view: a deferred step was launched since the last flush
  uint32_t* d_fc1_pending = nullptr;          // device view, read by the gated optimizer kernel
  cudaGraphExec_t graph_train_exec = nullptr;   // same step without the sampler (train on pre-sampled indexes)
  b200dqn_replay* graph_train_replay = nullptr;
  cudaStream_t graph_train_stream = nullptr;
  int graph_train_world = 0, graph_train_gen = 0;

  void* umma_state = nullptr;  // tcgen05 engine: fp16 operand planes + weight tile images (net_umma.cu)

  // multi-GPU
  void* nccl_comm = nullptr;
  int rank = 0, world = 1;
  // peer-memory gradient exchange (comm_p2p.cuh): d_g and its flag words as mapped from every rank
  bool xchg_ok = false;
  int xchg_flags = 0;             // k_xchg switches (comm_p2p.cuh), B200DQN_XCHG_FLAGS
  int xchg_blocks = 0;            // CTA cap per exchange (0 = kXMaxBlocks), B200DQN_XCHG_BLOCKS
  int xchg_sched = 2;             // 2: gather fc1's operands + LL all-reduce of the small layers (default),
                                  // 1: two-shot exchange per layer, 0: two-shot, two collectives
  float* xg[8] = {};              // [rank] = d_g
  uint32_t* xflags[8] = {};       // [rank] = d_xflags (tail of the d_g allocation)
  void* xopened[8] = {};          // cudaIpcOpenMemHandle results to close
  uint32_t* d_xflags = nullptr;   // flag words written by the peers
  uint32_t* d_xepoch = nullptr;   // local epoch counters [channel][block]
  uint32_t* d_xerr = nullptr;     // sticky "a wait timed out" word
  // gather / LL exchange: a second peer-mapped allocation, sized at comm_init for the world
  //   [push flags 4 KB][LL lines: parity x source x lines][H3 gather: parity x (hi | lo)][dZ4 gather: same]
  uint8_t* d_xbuf = nullptr;
  uint8_t* xbuf[8] = {};          // [rank] = d_xbuf
  void* xbuf_opened[8] = {};
  int64_t x_ll_off = 0, x_ll_lines = 0;               // bytes; LL lines per (parity, source)
  int64_t x_h3_off = 0, x_h3_parity = 0, x_h3_lo = 0;   // bytes; bytes per parity; lo plane offset in elements
  int64_t x_dz_off = 0, x_dz_parity = 0, x_dz_lo = 0;
  int64_t x_dzll_off = 0, x_dzll_lines = 0;           // LL line area of the dZ4 gather: bytes; lines per (parity, source)
  uint32_t* d_xll_epoch = nullptr;   // [kXChannels] epochs, then [kXChannels] tickets (LL exchange)
  uint32_t* d_xpush_epoch = nullptr; // [kXPushChannels] epochs, then tickets (plane push)
};
