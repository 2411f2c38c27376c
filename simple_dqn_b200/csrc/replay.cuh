// replay.cuh — device-resident replay ring and state window (object layouts).
#pragma once
#include "common.cuh"

struct b200dqn_replay {
  int device = 0;
  int64_t size = 0;
  int h = 0, w = 0, hist = 0, batch = 0;
  int64_t frame_bytes = 0;
  // host mirror of the cursor (src/replay_memory.py:17-18)
  int64_t count = 0, current = 0;

  // HBM
  uint8_t* d_screens = nullptr;    // [size][h][w]
  uint8_t* d_actions = nullptr;    // [size]
  int64_t* d_rewards = nullptr;    // [size]
  uint8_t* d_terminals = nullptr;  // [size] 0/1
  int64_t* d_cursor = nullptr;     // {count, current} — read by the sampler, graph-safe
  uint32_t* d_mt = nullptr;        // 624 key words + position (CPython random.getstate()[1])
  int32_t* d_idx = nullptr;        // [batch] accepted indexes, acceptance order
  uint32_t* d_words = nullptr;     // [0] words consumed by the last sample() call, [1] running total
  // host-mapped mirror written by the sampler: [0] samplings completed (published last), [1] words of the last
  // one, [2] running total
  volatile uint32_t* h_words = nullptr;
  uint32_t samples_launched = 0;   // host count of sampler launches (= the sequence number the next wait expects)
  // pinned staging for asynchronous MT19937 state uploads (the host stream moved since the last sampling)
  static constexpr int kMtSlots = 4;
  uint32_t* h_mt = nullptr;        // [kMtSlots][640]
  cudaEvent_t mt_done[kMtSlots] = {};
  int mt_slot = 0;
  uint8_t* d_pre = nullptr;        // [batch][hist][h][w]
  uint8_t* d_post = nullptr;       // [batch][hist][h][w]
  uint8_t* d_mb_actions = nullptr;
  int64_t* d_mb_rewards = nullptr;
  uint8_t* d_mb_terminals = nullptr;

  // pinned staging slots (small host<->device landing pads)
  static constexpr int kSlots = 16;
  uint8_t* h_stage = nullptr;
  cudaEvent_t slot_done[kSlots] = {};
  int next_slot = 0;
  bool rng_set = false;

  // add() is deferred: frames + scalars collect in one of two pinned banks and reach HBM in ONE
  // transfer + one tiny kernel when somebody needs the ring (sample / gather / getState / train) or
  // the bank is full.  The host cursor mirror is always current.
  static constexpr int kPend = 8;
  uint8_t* h_bank[2] = {};            // [kPend][frame_bytes] frames, then the scalar arrays
  cudaEvent_t bank_done[2] = {};
  int bank = 0, npend = 0;
  int64_t pend_pos0 = 0;              // ring slot of the first pending frame
  int64_t* bank_rewards(int b) const { return reinterpret_cast<int64_t*>(h_bank[b] + size_t(kPend) * frame_bytes); }
  uint8_t* bank_actions(int b) const { return reinterpret_cast<uint8_t*>(bank_rewards(b) + kPend); }
  uint8_t* bank_terminals(int b) const { return bank_actions(b) + kPend; }
};

struct b200dqn_statebuf {
  int device = 0;
  int h = 0, w = 0, hist = 0, batch = 0;
  int64_t frame_bytes = 0;
  uint8_t* d_buf = nullptr;  // [batch][hist][h][w]; only row 0 ever non-zero
  static constexpr int kSlots = 16;
  uint8_t* h_stage = nullptr;
  cudaEvent_t slot_done[kSlots] = {};
  int next_slot = 0;
};

namespace b200 {
// Launches used by the fused train step (net.cu).
int launch_sample(b200dqn_replay* r, cudaStream_t st);
// push the pending add()s to HBM (no-op when there are none); call before anything reads the ring
int replay_flush(b200dqn_replay* r, cudaStream_t st);
int replay_wait_words(b200dqn_replay* r, cudaStream_t st);
int replay_publish_words(b200dqn_replay* r, cudaStream_t st);   // device counters -> host-mapped mirror (tiny kernel)
#ifdef __CUDACC__
// [0] samplings completed (published last), [1] words of the last one, [2] running total
__device__ __forceinline__ void publish_words(const uint32_t* __restrict__ words, volatile uint32_t* host_words) {
  host_words[1] = words[0];
  host_words[2] = words[1];
  __threadfence_system();
  host_words[0] = words[2];
}
#endif
// adopt a host MT19937 state (624 key words + position) without synchronising the stream
int replay_set_rng_async(b200dqn_replay* r, const uint32_t* key624, uint32_t pos, cudaStream_t st);
}  // namespace b200
