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
  uint8_t* d_pre = nullptr;        // [batch][hist][h][w]
  uint8_t* d_post = nullptr;       // [batch][hist][h][w]
  uint8_t* d_mb_actions = nullptr;
  int64_t* d_mb_rewards = nullptr;
  uint8_t* d_mb_terminals = nullptr;

  // pinned staging slots for add()
  static constexpr int kSlots = 16;
  uint8_t* h_stage = nullptr;
  cudaEvent_t slot_done[kSlots] = {};
  int next_slot = 0;
  bool rng_set = false;
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
}  // namespace b200
