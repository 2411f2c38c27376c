// comm_p2p.cuh — gradient exchange between data-parallel learners through NVLink peer memory.
//
// Every rank maps every other rank's gradient buffer (cudaIpc handles, exchanged once at comm_init) and
// one kernel does the whole all-reduce of a parameter range in place, two-shot:
//
//   start barrier   block b of rank r tells block b of every peer "my gradient is complete" and waits
//                   for the same word from each of them (flags live in the receiver's memory, so the
//                   wait polls local HBM/L2);
//   reduce-scatter  rank r owns slice r: loads it from every rank over NVLink, adds in rank order
//                   0..W-1 (one rank computes each element once => all replicas get the same bits);
//   all-gather      and stores the sum straight into slice r of every rank's buffer;
//   end barrier     "my stores are done" to every peer, wait for theirs.  When the kernel retires,
//                   every slice of the local buffer holds the global sum.
//
// Barriers are per block index (no grid-wide sync, no co-residency requirement beyond in-order block
// dispatch), epochs are monotonic counters kept in device memory so the kernel replays inside a CUDA graph
// with constant arguments, and each call site owns a channel (its own flag words), so exchanges of
// different layers may run concurrently on different streams.  A wait that lasts 60 s sets a sticky error
// word and every later wait returns at once: a lost peer costs a bounded stall, never a hung GPU.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kXMaxWorld = 8;      // ranks addressable by one exchange (one NVSwitch domain)
constexpr int kXMaxBlocks = 128;   // CTAs per exchange kernel (= flag rows per channel)
constexpr int kXChannels = 6;      // concurrent call sites
constexpr int kXThreads = 256;
constexpr int kXUnroll = 2;        // float4 per thread per rank in flight
constexpr unsigned long long kXTimeoutNs = 60ull * 1000 * 1000 * 1000;   // bound of every device-side wait
constexpr int kXFlagWords = kXChannels * 2 * kXMaxBlocks * kXMaxWorld;

// ---- the default schedule ("gather", csrc/net.cu::backward_and_update_gather) moves far fewer bytes:
//  * fc1 (95 % of the parameters): dW4 = H3^T x dZ4 is NOT reduced.  Every rank pushes its 32 rows of the fp16
//    hi/lo planes of H3 (401 KB, ready after conv3_fwd) and dZ4 (65 KB, ready after the head) into every
//    rank's gather buffer (k_xpush: P2P stores + one release flag per peer), and each rank runs fc1_wgrad over
//    all W x 32 rows itself — W x 0.47 MB received instead of 2 x 6.4 MB moved, identical bits on every rank.
//  * conv1..3 and fc2 (1.3 MB together): one-shot all-reduce in NCCL's LL style (k_xll): every 16-byte line
//    carries 8 bytes of data and two copies of the epoch, pushed to every peer with one volatile 16-byte store;
//    the receiver polls its own memory until both flags match and adds the W contributions in rank order.
//    No barrier, no fence: latency is one NVLink store.  Receive buffers are double-buffered by epoch parity.
constexpr int kXPushChannels = 2;   // 0: H3 planes, 1: dZ4 planes

struct XllArgs {
  uint4* recv[kXMaxWorld];   // rank p's LL receive area as mapped here ([rank] = local)
  float4* g;                 // local gradient buffer (in: this rank's sum over its samples, out: global sum)
  int rank, world, chan;
  int64_t off4, ll4, n4;     // float4 offset in g, first LL element, count
  int64_t lines_per_src;     // LL lines per (parity, source rank)
  uint32_t* epoch;           // [kXChannels] completed exchanges per channel
  uint32_t* ticket;          // [kXChannels]
  uint32_t* err;
};

struct XPushArgs {
  const uint4* src[2];       // local hi plane, lo plane
  int64_t n16;               // 16-byte units per plane (this rank's rows)
  int64_t dst16[2];          // offset of this rank's rows inside one parity's gather area, per plane
  int64_t parity16;          // size of one parity's gather area
  uint4* gat[kXMaxWorld];    // rank p's gather area (parity 0) as mapped here
  uint32_t* pflags[kXMaxWorld];   // rank p's push flags [kXPushChannels][kXMaxWorld]
  int rank, world, chan;
  uint32_t* epoch;           // [kXPushChannels] completed pushes
  uint32_t* ticket;          // [kXPushChannels]
};

// dZ4 rows pushed by the head kernel itself (one CTA per sample): no separate push kernel between the head and the
// peers' fc1_wgrad.  Arrival is counted: every CTA adds 1 (red.release.sys) to counter [source rank] in each peer's
// exchange buffer after its stores; the consumer waits for rows x (pushes so far + 1).
constexpr int kXCountWord = 32;    // arrival counters [kXMaxWorld] start at this word of the exchange buffer
struct HeadPush {
  int world;                       // 0 = off
  int rank, rows;                  // rows per rank
  uint4* gat[kXMaxWorld];          // rank p's dZ4 gather area (parity 0) as mapped here
  uint32_t* cnt[kXMaxWorld];       // rank p's arrival counters as mapped here
  int64_t parity16, lo16;          // 16-byte units: one parity's area, offset of the lo plane inside it
  const uint32_t* epoch;           // completed dZ4 pushes (the next one uses parity (epoch + 1) & 1)
};

// dZ4 rows gathered in the LL protocol (k_xgather_ll): no flag word, no system-scope fence on the critical branch
// (the fence of the plain push costs ~8 us on this part) — every 16-byte unit of this rank's hi/lo rows travels as two
// LL lines to every peer, the receiver polls its own memory and writes the plain planes fc1_wgrad gathers from.
struct XGatherLL {
  const uint4* src[2];       // local hi plane, lo plane (this rank's rows)
  int64_t n16;               // 16-byte units per plane
  uint4* recv[kXMaxWorld];   // rank p's LL line area as mapped here
  uint4* gather;             // LOCAL plain gather area, parity 0
  int64_t parity16, lo16;    // plain area: units per parity, offset of the lo plane
  int64_t lines_per_src;     // LL lines per (parity, source) = 4 * n16
  int rank, world;
  uint32_t* epoch;           // completed dZ4 gathers (shared with the plain push channel 1)
  uint32_t* ticket;
  uint32_t* err;
  // the H3 rows (plain push, channel 0) must have landed as well before fc1_wgrad may start: their flags are polled
  // here, so that no separate wait kernel sits between this one and the consumer
  const uint32_t* h3_flags;  // [kXMaxWorld] local push flags of channel 0 (nullptr: do not wait)
  const uint32_t* h3_epoch;  // completed H3 pushes of this rank (= the epoch every peer's flag must have reached)
};

struct XPeers {
  float4* g[kXMaxWorld];        // rank p's gradient buffer as mapped here ([rank] = the local one)
  uint32_t* flags[kXMaxWorld];  // rank p's flag words
};

#ifdef __CUDACC__
__device__ __forceinline__ int xflag(int chan, int phase, int block, int src) {
  return ((chan * 2 + phase) * kXMaxBlocks + block) * kXMaxWorld + src;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_sys_f4(const float4* p) {   // strong load: never served from a stale L1 line
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_f4(float4* p, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// wait until *flag has reached epoch e (wrap-safe); false on the sticky error / timeout.  Polls with relaxed
// loads (an acquire per iteration would invalidate the SM's L1 under the co-resident kernels) and fences once.
__device__ __forceinline__ bool xwait(const uint32_t* flag, uint32_t e, uint32_t* err) {
  unsigned long long t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    if (int32_t(ld_relaxed_sys(flag) - e) >= 0) {
      asm volatile("fence.acq_rel.sys;" ::: "memory");
      return true;
    }
    __nanosleep(spins < 8 ? 40 : 200);   // back off: thousands of polling threads otherwise saturate the L2 the step's
                                         // own kernels live in (measured at W = 8: 20 us exchange kernels)
    if ((spins & 63) == 63) {
      if (*reinterpret_cast<volatile uint32_t*>(err)) return false;
      const unsigned long long now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > kXTimeoutNs) {
        atomicExch(err, 1u);
        return false;
      }
    }
  }
}

__device__ __forceinline__ void st_ll(uint4* p, uint32_t a, uint32_t b, uint32_t e) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(e), "r"(b), "r"(e) : "memory");
}
__device__ __forceinline__ uint4 ld_ll(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
// poll one line until both flag words carry epoch e; returns {data1, data2}
__device__ __forceinline__ bool ll_wait(const uint4* line, uint32_t e, uint32_t* err, float& a, float& b) {
  unsigned long long t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    const uint4 v = ld_ll(line);
    if (v.y == e && v.w == e) {
      a = __uint_as_float(v.x);
      b = __uint_as_float(v.z);
      return true;
    }
    __nanosleep(spins < 8 ? 40 : 200);
    if ((spins & 63) == 63) {
      if (*reinterpret_cast<volatile uint32_t*>(err)) break;
      const unsigned long long now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > kXTimeoutNs) {
        atomicExch(err, 1u);
        break;
      }
    }
  }
  a = b = 0.f;
  return false;
}

// raw variant: both data words as they are
__device__ __forceinline__ bool ll_wait_u(const uint4* line, uint32_t e, uint32_t* err, uint32_t& a, uint32_t& b) {
  float fa, fb;
  const bool ok = ll_wait(line, e, err, fa, fb);
  a = __float_as_uint(fa);
  b = __float_as_uint(fb);
  return ok;
}

#endif  // __CUDACC__

#if defined(__CUDACC__) && defined(B200_COMM_P2P_KERNELS)   // comm.cu owns the kernels
__device__ __forceinline__ float4 ld_weak_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_weak_f4(float4* p, const float4& v) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Tuning / diagnosis switches of k_xchg (b200dqn_debug_xchg measures them; production uses kXDefaultFlags)
constexpr int kXStrongLoads = 1;    // data loads ld.relaxed.sys instead of weak loads behind the acquire fence
constexpr int kXStrongStores = 2;   // data stores st.relaxed.sys instead of weak stores ahead of the release
constexpr int kXNoBarriers = 4;     // diagnosis only: skip both barriers (results undefined)
constexpr int kXNoData = 8;         // diagnosis only: barriers without the reduction

// all-reduce (sum) of float4 elements [off4, off4 + n4) of every rank's buffer, in place
__global__ void __launch_bounds__(kXThreads) k_xchg(XPeers pp, int rank, int world, int chan, int64_t off4, int64_t n4,
                                                     uint32_t* epochs, uint32_t* err, int flags, KTrace kt) {
  kt_begin(kt);
  const int b = blockIdx.x, t = threadIdx.x;
  __shared__ uint32_t s_epoch;
  if (t == 0) s_epoch = epochs[chan * kXMaxBlocks + b] + 1;
  __syncthreads();
  const uint32_t e = s_epoch;
  const bool barriers = !(flags & kXNoBarriers);
  // the producer kernel's writes are complete (stream order); the release publishes them system-wide
  if (barriers && t < world) {
    st_release_sys(pp.flags[t] + xflag(chan, 0, b, rank), e);
    xwait(pp.flags[rank] + xflag(chan, 0, b, t), e, err);
  }
  __syncthreads();

  const int64_t chunk = (n4 + world - 1) / world;
  const int64_t lo = rank * chunk, hi = min(n4, lo + chunk);
  const int64_t stride = int64_t(gridDim.x) * kXThreads;
  const bool sl = flags & kXStrongLoads, ss = flags & kXStrongStores;
  if (!(flags & kXNoData))
    for (int64_t i = lo + int64_t(b) * kXThreads + t; i < hi; i += stride * kXUnroll) {
      float4 v[kXUnroll][kXMaxWorld];
#pragma unroll
      for (int u = 0; u < kXUnroll; ++u)
#pragma unroll
        for (int p = 0; p < kXMaxWorld; ++p)
          if (p < world && i + u * stride < hi) {
            const float4* src = pp.g[p] + off4 + i + u * stride;
            v[u][p] = sl ? ld_sys_f4(src) : ld_weak_f4(src);
          }
#pragma unroll
      for (int u = 0; u < kXUnroll; ++u) {
        if (i + u * stride >= hi) break;
        float4 a = v[u][0];
#pragma unroll
        for (int p = 1; p < kXMaxWorld; ++p)
          if (p < world) {
            a.x += v[u][p].x;
            a.y += v[u][p].y;
            a.z += v[u][p].z;
            a.w += v[u][p].w;
          }
#pragma unroll
        for (int p = 0; p < kXMaxWorld; ++p)
          if (p < world) {
            float4* dst = pp.g[p] + off4 + i + u * stride;
            if (ss) st_sys_f4(dst, a);
            else st_weak_f4(dst, a);
          }
      }
    }

  __syncthreads();   // every thread's stores precede the release below (cumulativity through the barrier)
  if (barriers && t < world) {
    st_release_sys(pp.flags[t] + xflag(chan, 1, b, rank), e);
    xwait(pp.flags[rank] + xflag(chan, 1, b, t), e, err);
  }
  if (t == 0) epochs[chan * kXMaxBlocks + b] = e;
  kt_end(kt);
}

// ---- one-shot LL all-reduce ------------------------------------------------------------------------------
__global__ void __launch_bounds__(kXThreads) k_xll(XllArgs a, KTrace kt) {
  kt_begin(kt);
  const int t = threadIdx.x;
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(a.epoch + a.chan) + 1;   // same value in every block
  const int64_t par_base = int64_t(e & 1) * a.world * a.lines_per_src;
  const int64_t stride = int64_t(gridDim.x) * kXThreads;
  // push: this rank's values, two lines per float4, into slot [parity][rank] of every peer
  for (int64_t i = int64_t(blockIdx.x) * kXThreads + t; i < a.n4; i += stride) {
    const float4 v = a.g[a.off4 + i];
    const int64_t line = par_base + int64_t(a.rank) * a.lines_per_src + (a.ll4 + i) * 2;
#pragma unroll
    for (int p = 0; p < kXMaxWorld; ++p)
      if (p < a.world && p != a.rank) {
        st_ll(a.recv[p] + line, __float_as_uint(v.x), __float_as_uint(v.y), e);
        st_ll(a.recv[p] + line + 1, __float_as_uint(v.z), __float_as_uint(v.w), e);
      }
  }
  // collect: poll the local slots of the other ranks, add in rank order 0..W-1
  for (int64_t i = int64_t(blockIdx.x) * kXThreads + t; i < a.n4; i += stride) {
    const float4 mine = a.g[a.off4 + i];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    uint4 l0[kXMaxWorld], l1[kXMaxWorld];     // first pass: all peers' lines in flight at once
#pragma unroll
    for (int p = 0; p < kXMaxWorld; ++p)
      if (p < a.world && p != a.rank) {
        const uint4* line = a.recv[a.rank] + par_base + int64_t(p) * a.lines_per_src + (a.ll4 + i) * 2;
        l0[p] = ld_ll(line);
        l1[p] = ld_ll(line + 1);
      }
#pragma unroll
    for (int p = 0; p < kXMaxWorld; ++p) {
      if (p >= a.world) break;
      float4 v = mine;
      if (p != a.rank) {
        const uint4* line = a.recv[a.rank] + par_base + int64_t(p) * a.lines_per_src + (a.ll4 + i) * 2;
        if (l0[p].y == e && l0[p].w == e) { v.x = __uint_as_float(l0[p].x); v.y = __uint_as_float(l0[p].z); }
        else ll_wait(line, e, a.err, v.x, v.y);
        if (l1[p].y == e && l1[p].w == e) { v.z = __uint_as_float(l1[p].x); v.w = __uint_as_float(l1[p].z); }
        else ll_wait(line + 1, e, a.err, v.z, v.w);
      }
      if (p == 0) acc = v;
      else {
        acc.x += v.x;
        acc.y += v.y;
        acc.z += v.z;
        acc.w += v.w;
      }
    }
    a.g[a.off4 + i] = acc;
  }
  // the last block to finish publishes the new epoch (every block has read the old one by then)
  __syncthreads();
  if (t == 0) {
    __threadfence();
    if (atomicAdd(a.ticket + a.chan, 1u) == gridDim.x - 1) {
      a.ticket[a.chan] = 0;
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(a.epoch + a.chan) = e;
    }
  }
  kt_end(kt);
}

// ---- plane push: this rank's rows of two fp16 planes into every rank's gather area ------------------------
__global__ void __launch_bounds__(kXThreads) k_xpush(XPushArgs a, KTrace kt) {
  kt_begin(kt);
  const int t = threadIdx.x;
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(a.epoch + a.chan) + 1;
  const int64_t par = int64_t(e & 1) * a.parity16;
  const int64_t stride = int64_t(gridDim.x) * kXThreads;
  for (int64_t i = int64_t(blockIdx.x) * kXThreads + t; i < 2 * a.n16; i += stride) {
    const int pl = i >= a.n16 ? 1 : 0;
    const int64_t j = i - pl * a.n16;
    const uint4 v = a.src[pl][j];
#pragma unroll
    for (int p = 0; p < kXMaxWorld; ++p)
      if (p < a.world) a.gat[p][par + a.dst16[pl] + j] = v;
  }
  __syncthreads();
  __shared__ bool s_last;
  if (t == 0) {
    asm volatile("fence.acq_rel.sys;" ::: "memory");   // this block's stores (observed through the barrier) first
    s_last = atomicAdd(a.ticket + a.chan, 1u) == gridDim.x - 1;
    if (s_last) {
      asm volatile("fence.acq_rel.sys;" ::: "memory");
      a.ticket[a.chan] = 0;
      *reinterpret_cast<volatile uint32_t*>(a.epoch + a.chan) = e;
    }
  }
  __syncthreads();
  if (s_last && t < a.world) st_release_sys(a.pflags[t] + a.chan * kXMaxWorld + a.rank, e);
  kt_end(kt);
}

// ---- LL all-gather of a hi/lo plane pair (dZ4): push own units as LL lines, poll the peers', write plain planes
__global__ void __launch_bounds__(kXThreads) k_xgather_ll(XGatherLL a, KTrace kt) {
  kt_begin(kt);
  const int t = threadIdx.x;
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(a.epoch) + 1;
  const int64_t par_ll = int64_t(e & 1) * a.world * a.lines_per_src;
  uint4* plain = a.gather + int64_t(e & 1) * a.parity16;
  const int64_t stride = int64_t(gridDim.x) * kXThreads;
  for (int64_t i = int64_t(blockIdx.x) * kXThreads + t; i < 2 * a.n16; i += stride) {
    const int pl = i >= a.n16 ? 1 : 0;
    const int64_t j = i - pl * a.n16;
    const uint4 v = a.src[pl][j];
    const int64_t line = par_ll + int64_t(a.rank) * a.lines_per_src + 2 * i;
#pragma unroll
    for (int p = 0; p < kXMaxWorld; ++p)
      if (p < a.world && p != a.rank) {
        st_ll(a.recv[p] + line, v.x, v.y, e);
        st_ll(a.recv[p] + line + 1, v.z, v.w, e);
      }
    plain[pl * a.lo16 + int64_t(a.rank) * a.n16 + j] = v;     // this rank's own rows
  }
  for (int64_t i = int64_t(blockIdx.x) * kXThreads + t; i < 2 * a.n16; i += stride) {
    const int pl = i >= a.n16 ? 1 : 0;
    const int64_t j = i - pl * a.n16;
    // first pass: every peer's two lines requested together (one memory latency, not 2 (W - 1)); whatever has not
    // arrived yet is then waited for line by line
    uint4 l0[kXMaxWorld], l1[kXMaxWorld];
#pragma unroll
    for (int p = 0; p < kXMaxWorld; ++p)
      if (p < a.world && p != a.rank) {
        const uint4* line = a.recv[a.rank] + par_ll + int64_t(p) * a.lines_per_src + 2 * i;
        l0[p] = ld_ll(line);
        l1[p] = ld_ll(line + 1);
      }
#pragma unroll
    for (int p = 0; p < kXMaxWorld; ++p)
      if (p < a.world && p != a.rank) {
        const uint4* line = a.recv[a.rank] + par_ll + int64_t(p) * a.lines_per_src + 2 * i;
        uint4 v;
        if (l0[p].y == e && l0[p].w == e) { v.x = l0[p].x; v.y = l0[p].z; }
        else ll_wait_u(line, e, a.err, v.x, v.y);
        if (l1[p].y == e && l1[p].w == e) { v.z = l1[p].x; v.w = l1[p].z; }
        else ll_wait_u(line + 1, e, a.err, v.z, v.w);
        plain[pl * a.lo16 + int64_t(p) * a.n16 + j] = v;
      }
  }
  if (a.h3_flags && t < a.world) xwait(a.h3_flags + t, *reinterpret_cast<const volatile uint32_t*>(a.h3_epoch), a.err);
  __syncthreads();
  if (t == 0) {
    __threadfence();
    if (atomicAdd(a.ticket, 1u) == gridDim.x - 1) {
      *a.ticket = 0;
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(a.epoch) = e;
    }
  }
  kt_end(kt);
}

// wait until every rank's push of the current epoch has landed here (one block; runs ahead of the consumer).
// dz_rows > 0: the dZ4 rows come from the peers' head kernels and are COUNTED (HeadPush): wait for
// dz_rows x (count_epoch + 1) arrivals per source, then advance count_epoch and the dZ4 epoch (the parity selector).
__global__ void k_xwait(uint32_t* pflags, uint32_t* epoch, int world, uint32_t* err, int dz_rows, uint32_t* count_epoch,
                        KTrace kt) {
  kt_begin(kt);
  const int t = threadIdx.x;
  if (t < world) {
    xwait(pflags + t, *reinterpret_cast<const volatile uint32_t*>(epoch), err);            // channel 0: H3 planes
  } else if (t < 2 * world && dz_rows >= 0) {      // dz_rows < 0: the dZ4 rows do not come through this wait at all
    const int p = t - world;
    if (dz_rows == 0)
      xwait(pflags + kXMaxWorld + p, *reinterpret_cast<const volatile uint32_t*>(epoch + 1), err);
    else
      xwait(pflags + kXCountWord + p, uint32_t(dz_rows) * (*reinterpret_cast<const volatile uint32_t*>(count_epoch) + 1u), err);
  }
  __syncthreads();
  if (dz_rows > 0 && t == 0) {
    count_epoch[0] += 1;
    epoch[1] += 1;
  }
  kt_end(kt);
}
#endif  // B200_COMM_P2P_KERNELS

}  // namespace b200
