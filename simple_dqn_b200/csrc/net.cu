// net.cu — Nature-DQN train / predict on the device behind DeepQNetwork's call surface
// (src/deepqnetwork.py:15-192 of the reference; per-entry citations in include/b200dqn.h).
#include <stdlib.h>

#include <new>
#include <vector>

#include <cuda_fp16.h>

#include "net.cuh"
#include "net_umma.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------
// K2e + K3 + K4a/K5a ("head"): finish fc1 (sum split-K partials, Rectlin), run fc2 (Affine
// nout=A, no activation) for both networks of ONE sample per CTA, then — same CTA, no inter-CTA
// hand-off — the TD target / delta / cost / clip of src/deepqnetwork.py:124-159 and the fc2 backward:
//   dZ4[b][k] = (sum_a delta[b][a] W5[k][a]) * (H4[b][k] > 0),  dW5[k][a] = sum_b H4[b][k] delta[b][a].
// The reference forms the target on the host in Python floats (double) and stores it into a
// float32 array; we do the same arithmetic in fp64 and round once.  The per-sample cost
// 0.5*sum_a delta^2 (BEFORE the clip) goes to row_cost; the batch mean is formed off the critical
// chain by k_cost_finish.  grid = rows, block = 512 (one thread per hidden unit).
// ------------------------------------------------------------------------------------------
struct HeadTrainArgs {
  int enable;
  const uint8_t* actions;
  const int64_t* rewards;
  const uint8_t* terminals;
  const int32_t* midx;
  double discount;
  int min_reward, max_reward;
  float clip;
  float* delta;       // [rows][A]
  const uint32_t* step;   // completed train steps (k_cost_finish increments it)
  float* row_cost;    // [rows]
  float* dz4;         // [rows][512]
  float* dw5_rows;    // [rows][512][A] per-row partials of dW5 (summed in row order by the optimizer)
  __half* dz4_hi;     // fp16 hi / scaled-lo planes of dZ4 for the tcgen05 dgrad (nullptr in fp32 mode)
  int64_t dz4_lo_off;
  float* adam_l;      // Adam only: this step's scalar l = lr*sqrt(1-beta_2^t)/(1-beta_1^t) for the optimizer kernels
  float adam_lr;
  int num_actions;    // actions[] >= num_actions would index q / W5 out of bounds: flagged in err, clamped
  uint32_t* err;
  HeadPush push;      // data-parallel gather schedule: this CTA's dZ4 row goes straight to every rank (world = 0: off)
};

__global__ void __launch_bounds__(kHidden)
k_head(const float* __restrict__ part, int splits, int rows, int nets, float* h4_online, float* h4_target,
       const float* __restrict__ w5_online, const float* __restrict__ w5_target, float* q_online,
       float* q_target, int A, const HeadTrainArgs td, const KTrace kt) {
  __shared__ float red[2][kHidden / 32][kMaxActions];
  __shared__ float s_q[2][kMaxActions];
  __shared__ float s_d;
  __shared__ int s_a;
  __shared__ __align__(16) __half s_row[2][kHidden];   // hi / lo of this sample's dZ4 row (peer push)
  const int b = blockIdx.x, t = threadIdx.x;
  kt_begin(kt);
  // The TD scalars of this sample depend only on the sampler (several kernels upstream, complete by now):
  // fetch them before the dependency wait, off the tail of the kernel.
  int td_a = 0, td_term = 0;
  int64_t td_r = 0;
  if (td.enable && t == 0) {
    const int64_t mi = td.midx[b];
    td_a = td.actions[mi];
    td_r = td.rewards[mi];
    td_term = td.terminals[mi];
    if (td_a >= td.num_actions) {   // the reference would raise IndexError (deepqnetwork.py:141); here: sticky flag
      atomicExch(td.err, 1u);
      td_a = td.num_actions - 1;
    }
  }
  if (td.enable && td.adam_l && b == 0 && t == 32) {
    // Adam.optimize: self.t += 1;  l = lr * sqrt(1 - beta_2**t) / (1 - beta_1**t)  (Python doubles, fp32 tensor ops)
    const double tt = double(*td.step) + 1.0;
    const float a = float(1.0 - pow(0.999, tt)), c = float(1.0 - pow(0.9, tt));
    *td.adam_l = __fdiv_rn(__fmul_rn(td.adam_lr, __fsqrt_rn(a)), c);
  }
  pdl_wait();
  pdl_launch_dependents();
  float h[2] = {0.f, 0.f};
#pragma unroll
  for (int z = 0; z < 2; ++z) {
    if (z < nets) {
      float acc = 0.f;
      for (int s = 0; s < splits; ++s) acc += part[((z * splits + s) * rows + b) * kHidden + t];
      h[z] = fmaxf(acc, 0.f);
      (z ? h4_target : h4_online)[b * kHidden + t] = h[z];
    }
  }
#pragma unroll
  for (int z = 0; z < 2; ++z) {
    if (z < nets) {
      const float* w5 = z ? w5_target : w5_online;
      for (int a = 0; a < A; ++a) {
        float v = h[z] * w5[t * A + a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((t & 31) == 0) red[z][t >> 5][a] = v;
      }
    }
  }
  __syncthreads();
  if (t < nets * A) {
    const int z = t / A, a = t % A;
    float v = 0.f;
#pragma unroll
    for (int wI = 0; wI < kHidden / 32; ++wI) v += red[z][wI][a];
    (z ? q_target : q_online)[b * A + a] = v;
    s_q[z][a] = v;
  }
  if (!td.enable) {
    kt_end(kt);
    return;
  }
  __syncthreads();
  if (t == 0) {
    const int a = td_a;
    int64_t r = td_r;
    r = r < td.min_reward ? td.min_reward : (r > td.max_reward ? td.max_reward : r);     // np.clip (:136)
    float maxq = s_q[1][0];
    for (int j = 1; j < A; ++j) maxq = fmaxf(maxq, s_q[1][j]);                            // be.max(postq) (:124)
    const double y = td_term ? double(r) : double(r) + td.discount * double(maxq);          // :140-143
    const float target = static_cast<float>(y);
    float d = s_q[0][a] - target;                                                         // SumSquared grad (:149)
    td.row_cost[b] = 0.5f * d * d;                                                        // :154, before the clip
    if (td.clip > 0.f) d = fminf(fmaxf(d, -td.clip), td.clip);                            // :158-159
    for (int j = 0; j < A; ++j) td.delta[b * A + j] = (j == a) ? d : 0.f;
    s_d = d;
    s_a = a;
  }
  __syncthreads();
  {
    const float d = s_d;
    const int a = s_a;
    const float hv = h[0];
    const float o = hv > 0.f ? d * w5_online[t * A + a] : 0.f;      // delta is non-zero only at the taken action
    td.dz4[b * kHidden + t] = o;
    if (td.dz4_hi) {
      const __half hh = __float2half_rn(o);
      const __half ll = __float2half_rn((o - __half2float(hh)) * 2048.0f);
      td.dz4_hi[b * kHidden + t] = hh;
      td.dz4_hi[td.dz4_lo_off + b * kHidden + t] = ll;
      s_row[0][t] = hh;
      s_row[1][t] = ll;
    }
    float* dw = td.dw5_rows + (int64_t(b) * kHidden + t) * A;        // per-row partial, summed by the optimizer
    for (int j = 0; j < A; ++j) dw[j] = (j == a) ? hv * d : 0.f;
  }
  if (td.push.world > 0) {
    // this sample's dZ4 row (1 KB per plane) to every rank's gather area, 16 bytes per store, then one counted
    // arrival per peer: the peers' fc1_wgrad over the global minibatch needs nothing else from this rank
    __syncthreads();
    const HeadPush& hp = td.push;
    if (t < 2 * (kHidden / 8)) {
      const int pl = t / (kHidden / 8), c = t % (kHidden / 8);
      const uint4 v = reinterpret_cast<const uint4*>(s_row[pl])[c];
      const int64_t par = int64_t((*reinterpret_cast<const volatile uint32_t*>(hp.epoch) + 1u) & 1u) * hp.parity16;
      const int64_t at = par + pl * hp.lo16 + (int64_t(hp.rank) * hp.rows + b) * (kHidden / 8) + c;
#pragma unroll
      for (int p = 0; p < kXMaxWorld; ++p)
        if (p < hp.world) hp.gat[p][at] = v;
    }
    __syncthreads();
    if (t < hp.world) {
      asm volatile("fence.acq_rel.sys;" ::: "memory");   // the CTA's stores (observed through the barrier) first
      asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(hp.cnt[t] + hp.rank) : "memory");
    }
  }
  kt_end(kt);
}

// cost = mean over the batch of the per-sample costs (GeneralizedCost.get_cost, src/deepqnetwork.py:154), summed in
// row order by one thread (deterministic); advances the cost ring and the step counter.  Runs off the critical
// chain (the stream of the fc2 optimizer): nothing on the device waits for the scalar.
__global__ void __launch_bounds__(256)
k_cost_finish(const float* __restrict__ row_cost, int rows, float* cost_ring, uint32_t* step,
              volatile uint32_t* host_res, const uint32_t* __restrict__ sampler_words, volatile uint32_t* host_words,
              const KTrace kt) {
  __shared__ float s_c[1024];
  kt_begin(kt);
  float tot = 0.f;
  for (int base = 0; base < rows; base += 1024) {
    const int n = min(1024, rows - base);
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_c[i] = row_cost[base + i];
    __syncthreads();
    if (threadIdx.x == 0)
      for (int i = 0; i < n; ++i) tot += s_c[i];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const uint32_t sidx = *step;
    const float cost = tot / float(rows);
    cost_ring[sidx % kCostRing] = cost;
    cost_ring[kCostRing] = cost;   // "latest" slot
    *step = sidx + 1;
    if (host_res) {                // the host polls [0]: data first, system fence, then the sequence number
      host_res[4 + sidx % kHostCosts] = __float_as_uint(cost);
      host_res[1] = __float_as_uint(cost_ring[kCostRing + 1]);   // action-range flag word
      if (sampler_words) {         // this step's index draw: words consumed, for the host's lock-step `random`
        host_words[1] = sampler_words[0];
        host_words[2] = sampler_words[1];
      }
      __threadfence_system();
      if (sampler_words) host_words[0] = sampler_words[2];
      host_res[0] = sidx + 1;
    }
  }
  kt_end(kt);
}

// Small-layer optimizer without tile images (fc2: 512 x A parameters, CUDA-core layer): 8 lanes per float4 sum the
// split partials (lane l takes partials l, l+8, ... in order; fixed xor tree across lanes — the summation order of
// k_optimizer, bit-identical) and lane 0 applies the configured update.
__global__ void __launch_bounds__(256)
k_opt_small(const float* __restrict__ part, int splits, int64_t size, float* __restrict__ w, float* __restrict__ sst,
            const OptArgs opt, const KTrace kt) {
  kt_begin(kt);
  pdl_wait();
  pdl_launch_dependents();
  const int tid = threadIdx.x, lane8 = tid & 7;
  const int64_t i = (int64_t(blockIdx.x) * 32 + (tid >> 3)) * 4;
  const bool live = i < size;
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) {
    const float* p = part + i;
#pragma unroll 4
    for (int sp = lane8; sp < splits; sp += 8) {
      const float4 v = *reinterpret_cast<const float4*>(p + int64_t(sp) * size);
      g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
    }
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    g.x += __shfl_xor_sync(0xffffffffu, g.x, o);
    g.y += __shfl_xor_sync(0xffffffffu, g.y, o);
    g.z += __shfl_xor_sync(0xffffffffu, g.z, o);
    g.w += __shfl_xor_sync(0xffffffffu, g.w, o);
  }
  if (live && lane8 == 0) {
    float nw[4];
    opt_update_vec<4>(opt, opt_step_scalar(opt), reinterpret_cast<const float*>(&g), nw, w + i, sst + i);
  }
  kt_end(kt);
}

// ------------------------------------------------------------------------------------------
// K6: gradient reduction + the configured Neon optimizer (src/deepqnetwork.py:50-61,165; rules in optim.cuh).
// RMSProp:  g = dW / bsz;  s = decay*s + g*g*(1-decay);  W = W - (g*lr) / (sqrt(s + eps) + eps)
// The split-K partials of every layer are summed here in fixed order (deterministic), so the
// wgrad kernels never need atomics.  mode bit0: sum partials (else read g_buf); bit1: write the
// summed gradient to g_buf (all-reduce input / get_grads); bit2: apply the update.
// Explicit _rn intrinsics keep the compiler from contracting into FMAs, so given identical
// gradients the update is bit-identical to the numpy oracle.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_optimizer(const LayerTable lt, const float* __restrict__ part, float* __restrict__ g_buf, float* __restrict__ w,
            float* __restrict__ s, int64_t b4, int64_t e4, int mode, const OptArgs opt, const KTrace kt) {
  const int64_t i4 = b4 + blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  kt_begin(kt);
  pdl_wait();
  pdl_launch_dependents();
  if (i4 >= e4) return;
  const int64_t i = i4 * 4;
  float4 g;
  if (mode & 1) {
    int l = 0;
#pragma unroll
    for (int j = 1; j < kLayers; ++j) l += (i >= lt.off[j]) ? 1 : 0;
    const int64_t lsize = lt.off[l + 1] - lt.off[l];
    const float* p = part + lt.part_off[l] + (i - lt.off[l]);
    // Fixed summation tree shared with k_opt_conv (net_umma.cu): eight strided running sums
    // a[l] = sum_{sp = l (mod 8)} part[sp], combined as ((a0+a1)+(a2+a3)) + ((a4+a5)+(a6+a7)).
    const int nsp = lt.splits[l];
    float4 a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int sp0 = 0; sp0 < nsp; sp0 += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (sp0 + u < nsp) {
          const float4 v = *reinterpret_cast<const float4*>(p + (sp0 + u) * lsize);
          a[u].x += v.x; a[u].y += v.y; a[u].z += v.z; a[u].w += v.w;
        }
      }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1)
#pragma unroll
      for (int u = 0; u < 8; u += 2 * o) {
        a[u].x += a[u + o].x; a[u].y += a[u + o].y; a[u].z += a[u + o].z; a[u].w += a[u + o].w;
      }
    g = a[0];
  } else {
    g = *reinterpret_cast<const float4*>(g_buf + i);
  }
  if (mode & 2) *reinterpret_cast<float4*>(g_buf + i) = g;
  if (mode & 4) {
    float nw[4];
    opt_update_vec<4>(opt, opt_step_scalar(opt), reinterpret_cast<const float*>(&g), nw, w + i, s + i);
  }
  kt_end(kt);
}

__global__ void k_iota(int32_t* a, int32_t* b, int n, int mult) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    a[i] = i;
    b[i] = i * mult;
  }
}

__global__ void k_zero_rows(float* q, int from, int to, int A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (to - from) * A) q[from * A + i] = 0.f;
}

static inline unsigned cdiv(int64_t a, int64_t b) { return unsigned((a + b - 1) / b); }
static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

template <class P, int BM, int BN, int BK, int TM, int TN>
static int launch_gemm(const char* label, const P& p, int M, int N, int Z, cudaStream_t st) {
  dim3 grid(cdiv(M, BM), cdiv(N, BN), Z);
  ++g_launch_count;
  k_simt_gemm<P, BM, BN, BK, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, st>>>(p);
  B2_LAUNCH_CHECK();
  B2_PROF(label, st);
  return B200DQN_OK;
}

// Where the first conv layer reads its frames: in place from the ring (fused) or from staged states.
struct FrameSource {
  const uint8_t* src[2];
  const int32_t* idx[2];
  int shift[2];
  int64_t nframes[2];   // frames in each source array
};

// Model.fprop for `nets` networks (z = 0 online, z = 1 target) on `rows` samples.
static int forward(b200dqn_net* n, const FrameSource& fs, int nets, int rows, cudaStream_t st,
                   const HeadTrainArgs& td) {
  const LayerTable& lt = n->lt;
  const float* w[2] = {n->d_w, n->d_tw};
  int rc;
  if (n->cfg.math_mode == B200DQN_MATH_TCGEN05) {
    rc = umma_forward(n, fs.src, fs.idx, fs.shift, fs.nframes, nets, rows, st);
    if (rc) return rc;
  } else {
    {
      Conv1Fwd p;
      for (int z = 0; z < 2; ++z) {
        p.src[z] = fs.src[z]; p.idx[z] = fs.idx[z]; p.shift[z] = fs.shift[z];
        p.w[z] = w[z] + lt.off[0]; p.out[z] = n->d_h1[z];
      }
      p.nb = rows;
      if ((rc = launch_gemm<Conv1Fwd, 64, 32, 16, 4, 2>("conv1_fwd", p, rows * kP1 * kP1, kC1, nets, st))) return rc;
    }
    {
      using P = ConvFwd<kP1, kC1, 4, 2, kC2>;
      P p;
      for (int z = 0; z < 2; ++z) { p.in[z] = n->d_h1[z]; p.w[z] = w[z] + lt.off[1]; p.out[z] = n->d_h2[z]; }
      p.nb = rows;
      if ((rc = launch_gemm<P, 32, 64, 16, 2, 4>("conv2_fwd", p, rows * kP2 * kP2, kC2, nets, st))) return rc;
    }
    {
      using P = ConvFwd<kP2, kC2, 3, 1, kC3>;
      P p;
      for (int z = 0; z < 2; ++z) { p.in[z] = n->d_h2[z]; p.w[z] = w[z] + lt.off[2]; p.out[z] = n->d_h3[z]; }
      p.nb = rows;
      if ((rc = launch_gemm<P, 32, 64, 16, 2, 4>("conv3_fwd", p, rows * kP3 * kP3, kC3, nets, st))) return rc;
    }
    {
      Fc1Fwd p;
      for (int z = 0; z < 2; ++z) { p.in[z] = n->d_h3[z]; p.w[z] = w[z] + lt.off[3]; }
      p.part = n->d_fc1part; p.nb = rows; p.splits = kFc1Splits; p.kchunk = kFc1Chunk;
      if ((rc = launch_gemm<Fc1Fwd, 32, 64, 16, 2, 4>("fc1_fwd", p, rows, kHidden, nets * kFc1Splits, st))) return rc;
    }
  }
  const int fc1_splits = n->cfg.math_mode == B200DQN_MATH_TCGEN05 ? umma_fc1_splits(rows) : kFc1Splits;
  B2_CHECK_CUDA(launch_pdl(k_head, dim3(rows), dim3(kHidden), 0, st, (const float*)n->d_fc1part, fc1_splits, rows, nets,
                           n->d_h4[0], n->d_h4[1], w[0] + lt.off[4], w[1] + lt.off[4], n->d_q[0], n->d_q[1], n->A,
                           td, ktrace_slot("head")));
  B2_PROF(td.enable ? "head(fc2+td+fc2_bwd)" : "fc2_fwd", st);
  return B200DQN_OK;
}

static int wgrad_chunk(int kred, int base) {
  int c = (kred + 31) / 32;
  if (c < base) c = base;
  return round_up(c, 16);
}

enum BwdOp { kFc1Wgrad, kFc1Dgrad, kConv3Wgrad, kConv3Dgrad, kConv2Wgrad, kConv2Dgrad, kConv1Wgrad };

// One GEMM-shaped backward op on stream `st`, on whichever engine math_mode selects.
static int bwd_op(b200dqn_net* n, const FrameSource& fs, int rows, BwdOp op, cudaStream_t st) {
  const LayerTable& lt = n->lt;
  const float* w = n->d_w;
  if (n->cfg.math_mode == B200DQN_MATH_TCGEN05 && umma_has_backward())
    return umma_backward_op(n, int(op), fs.src[0], fs.idx[0], fs.shift[0], rows, st);
  switch (op) {
    case kFc1Wgrad: {
      Fc1Wgrad p{n->d_h3[0], n->d_dz4, n->d_part + lt.part_off[3], rows};
      return launch_gemm<Fc1Wgrad, 64, 64, 16, 4, 4>("fc1_wgrad", p, kFlat, kHidden, 1, st);
    }
    case kFc1Dgrad: {
      Fc1Dgrad p{n->d_dz4, w + lt.off[3], n->d_h3[0], n->d_dz3, rows};
      return launch_gemm<Fc1Dgrad, 32, 32, 16, 2, 2>("fc1_dgrad", p, rows, kFlat, 1, st);
    }
    case kConv3Wgrad: {
      using P = ConvWgrad<kP2, kC2, 3, 1, kC3>;
      P p{n->d_h2[0], n->d_dz3, n->d_part + lt.part_off[2], rows, wgrad_chunk(rows * kP3 * kP3, 112)};
      return launch_gemm<P, 64, 64, 16, 4, 4>("conv3_wgrad", p, P::KW, kC3, lt.splits[2], st);
    }
    case kConv3Dgrad: {
      using P = ConvDgrad<kP2, kC2, 3, 1, kC3>;
      P p{n->d_dz3, w + lt.off[2], n->d_h2[0], n->d_dz2, rows};
      return launch_gemm<P, 32, 64, 16, 2, 4>("conv3_dgrad", p, rows * P::HC * P::HC, kC2, 1, st);
    }
    case kConv2Wgrad: {
      using P = ConvWgrad<kP1, kC1, 4, 2, kC2>;
      P p{n->d_h1[0], n->d_dz2, n->d_part + lt.part_off[1], rows, wgrad_chunk(rows * kP2 * kP2, 96)};
      return launch_gemm<P, 64, 64, 16, 4, 4>("conv2_wgrad", p, P::KW, kC2, lt.splits[1], st);
    }
    case kConv2Dgrad: {
      using P = ConvDgrad<kP1, kC1, 4, 2, kC2>;
      P p{n->d_dz2, w + lt.off[1], n->d_h1[0], n->d_dz1, rows};
      return launch_gemm<P, 64, 32, 16, 4, 2>("conv2_dgrad", p, rows * P::HC * P::HC, kC1, 4, st);
    }
    default: {
      Conv1Wgrad p{fs.src[0], fs.idx[0], fs.shift[0], n->d_dz1, n->d_part + lt.part_off[0], rows,
                   wgrad_chunk(rows * kP1 * kP1, 512)};
      return launch_gemm<Conv1Wgrad, 64, 32, 16, 4, 2>("conv1_wgrad", p, kK1, kC1, lt.splits[0], st);
    }
  }
}

// RMSProp (or gradient reduction) over layers [l0, l1] on stream st; mode bits as in k_optimizer.
// The optimizer constants of this net (src/deepqnetwork.py:50-59; Neon's defaults for what the reference leaves unset)
OptArgs make_opt_args(const b200dqn_net* n, int rows) {
  OptArgs o{};
  o.kind = n->cfg.optimizer;
  o.nstates = n->n_states;
  o.bsz = float(rows * n->world);
  o.lr = float(n->cfg.learning_rate);
  o.decay = float(n->cfg.decay_rate);
  o.one_m_decay = float(1.0 - n->cfg.decay_rate);
  o.eps = 1e-6f;
  o.b1 = float(0.9); o.one_m_b1 = float(1.0 - 0.9);
  o.b2 = float(0.999); o.one_m_b2 = float(1.0 - 0.999);
  o.adam_eps = 1e-8f;
  o.adam_l = n->d_optscal;
  o.plane = n->n_params;
  return o;
}

// optimizer update (or gradient reduction) over layers [l0, l1] on stream st; mode bits as in k_optimizer.
static int optimizer_range(b200dqn_net* n, int l0, int l1, int mode, int rows, cudaStream_t st, const char* label) {
  const LayerTable& lt = n->lt;
  const int64_t b4 = lt.off[l0] / 4, e4 = lt.off[l1 + 1] / 4;
  B2_CHECK_CUDA(launch_pdl(k_optimizer, dim3(cdiv(e4 - b4, 256)), dim3(256), 0, st, lt, (const float*)n->d_part, n->d_g,
                           n->d_w, n->d_s, b4, e4, mode, make_opt_args(n, rows), ktrace_slot(label)));
  B2_PROF(label, st);
  if (mode & 4) return umma_pack_layers(n, 0, l0, l1, st);   // refresh the fp16 tile images of the updated layers
  return B200DQN_OK;
}

// batch-mean cost -> cost ring, step counter + 1 (once per train step, after the head, on any stream behind it)
static int cost_finish_on(b200dqn_net* n, int rows, cudaStream_t s) {
  NoPdlScope plain;
  b200dqn_replay* r = n->step_replay;     // the ring this step samples from (nullptr: host-supplied minibatch)
  B2_CHECK_CUDA(launch_pdl(k_cost_finish, dim3(1), dim3(256), 0, s, (const float*)n->d_rowcost, rows, n->d_cost, n->d_step,
                           n->h_res, (const uint32_t*)(r ? r->d_words : nullptr),
                           (volatile uint32_t*)(r ? r->h_words : nullptr), ktrace_slot("cost")));
  B2_PROF("cost", s);
  return B200DQN_OK;
}

// fc2 update from the head's per-row partials (single-GPU schedules): 8-lane reduction, no image
static int opt_fc2_small(b200dqn_net* n, int rows, cudaStream_t s) {
  const LayerTable& lt = n->lt;
  const int64_t size = lt.off[5] - lt.off[4];
  B2_CHECK_CUDA(launch_pdl(k_opt_small, dim3(cdiv(size / 4, 32)), dim3(256), 0, s, (const float*)n->d_part + lt.part_off[4],
                           lt.splits[4], size, n->d_w + lt.off[4], n->d_s + lt.off[4], make_opt_args(n, rows),
                           ktrace_slot("opt_fc2")));
  B2_PROF("opt_fc2", s);
  return B200DQN_OK;
}

#define B2_TRY(expr)          \
  do {                        \
    int rc__ = (expr);        \
    if (rc__) return rc__;    \
  } while (0)

// Model.bprop + optimizer.optimize for the online network (src/deepqnetwork.py:162-165).
//
// The dgrad chain fc1 -> conv3 -> conv2 is the critical path; every wgrad only needs the dZ of its
// own layer, and every per-layer RMSProp update only needs that layer's wgrad plus the guarantee
// that the dgrad reading the old weights has finished.  On a single GPU those independent pieces
// run on three side streams (graph branches under capture):
//   main : head . fc1_dgrad . conv3_dgrad . conv2_dgrad . conv1_wgrad . opt(conv1)
//   sA   :          fc1_wgrad ......... [after fc1_dgrad]   opt(fc1, fc2)
//   sB   :                    conv3_wgrad .. [after conv3_dgrad] opt(conv3)
//   sC   :                               conv2_wgrad .. [after conv2_dgrad] opt(conv2)
// In a communicator the update follows one all-reduce of the whole gradient, so the simple
// serial order is kept.
// Data-parallel schedule (communicator, tcgen05 engine): the fc gradient (95 % of the bytes) is summed and
// all-reduced as soon as fc1_wgrad is done, hidden behind the dgrad chain; the three small conv gradients
// share one all-reduce at the tail.  (Measured on 2x B200: one collective at the tail 190 us/step, this
// two-collective schedule 144 us/step, one collective per layer 166 us/step — small NCCL all-reduces
// cost ~15-20 us each inside the graph, so fewer is better once the big one is hidden.)  Both collectives run in this order on one dedicated stream (a NCCL
// communicator must not be used from two streams at once); updates read the reduced gradient from d_g.
// Peer-memory exchange (comm_p2p.cuh): no shared communicator, so every layer's gradient is reduced
// the moment its wgrad has finished, on that layer's own branch, and only conv1's 32 KB exchange is left
// on the critical chain:  wgrad -> partial sums -> exchange (in place, all ranks) -> RMSProp from d_g.
// ---- software-pipelined fc1 update (net.cuh: graph_def_exec) ----------------------------------------------------
// In the deferred variant of the step graph the fc1 optimizer does not run where dW4 becomes available; the step
// only marks the update as pending ...
static int fc1_update_deferred(b200dqn_net* n, cudaStream_t branch) {
  B2_CHECK_CUDA(cudaMemsetAsync(n->d_fc1_pending, 1, sizeof(uint32_t), branch));   // != 0
  return B200DQN_OK;
}
// ... the NEXT step applies it first thing, on a side branch under its forward convolutions (joined before fc1_fwd) ...
static int fc1_update_leading(b200dqn_net* n, cudaStream_t st) {
  cudaStream_t sA = n->side[0];
  B2_CHECK_CUDA(cudaEventRecord(n->ev[15], st));
  B2_CHECK_CUDA(cudaStreamWaitEvent(sA, n->ev[15], 0));
  {
    NoPdlScope side;
    B2_TRY(umma_opt_fc1(n, n->nb, sA, false, n->d_fc1_pending));
  }
  B2_CHECK_CUDA(cudaEventRecord(n->ev[16], sA));
  return B200DQN_OK;
}
// ... and train_fused applies the last one before it returns.
static int fc1_update_flush(b200dqn_net* n, cudaStream_t st) {
  if (!n->fc1_pending) return B200DQN_OK;
  {
    NoPdlScope plain;
    B2_TRY(umma_opt_fc1(n, n->nb, st, false, n->d_fc1_pending));
  }
  B2_CHECK_CUDA(cudaMemsetAsync(n->d_fc1_pending, 0, sizeof(uint32_t), st));
  n->fc1_pending = false;
  return B200DQN_OK;
}
static void destroy_step_graphs(b200dqn_net* n) {
  if (n->graph_exec) { cudaGraphExecDestroy(n->graph_exec); n->graph_exec = nullptr; }
  if (n->graph_def_exec) { cudaGraphExecDestroy(n->graph_def_exec); n->graph_def_exec = nullptr; }
}

static int backward_and_update_xchg(b200dqn_net* n, const FrameSource& fs, int rows, cudaStream_t st) {
  cudaStream_t sA = n->side[0], sB = n->side[1], sC = n->side[2];
  cudaEvent_t* ev = n->ev;
  B2_CHECK_CUDA(cudaEventRecord(ev[0], st));                 // head done: dZ4, dW5 partials
  B2_CHECK_CUDA(cudaStreamWaitEvent(sA, ev[0], 0));
  {
    NoPdlScope side;
    B2_TRY(bwd_op(n, fs, rows, kFc1Wgrad, sA));
    B2_TRY(cost_finish_on(n, rows, sA));
    B2_TRY(optimizer_range(n, 3, 4, 1 | 2, rows, sA, "reduce_fc"));
    B2_TRY(comm_xchg_range(n, 3, 4, 3, sA, "xchg_fc"));
  }
  B2_TRY(bwd_op(n, fs, rows, kFc1Dgrad, st));
  B2_CHECK_CUDA(cudaEventRecord(ev[1], st));                 // W4 no longer needed
  {
    NoPdlScope side;
    B2_CHECK_CUDA(cudaStreamWaitEvent(sA, ev[1], 0));
    B2_TRY(umma_opt_fc1(n, rows, sA, true));
    B2_TRY(optimizer_range(n, 4, 4, 4, rows, sA, "opt_fc2"));
    B2_CHECK_CUDA(cudaStreamWaitEvent(sB, ev[1], 0));
    B2_TRY(bwd_op(n, fs, rows, kConv3Wgrad, sB));
    B2_TRY(optimizer_range(n, 2, 2, 1 | 2, rows, sB, "reduce_conv3"));
    B2_TRY(comm_xchg_range(n, 2, 2, 2, sB, "xchg_conv3"));
  }
  B2_TRY(bwd_op(n, fs, rows, kConv3Dgrad, st));
  B2_CHECK_CUDA(cudaEventRecord(ev[2], st));                 // W3 no longer needed
  {
    NoPdlScope side;
    B2_CHECK_CUDA(cudaStreamWaitEvent(sB, ev[2], 0));
    B2_TRY(umma_opt_conv(n, 2, rows, sB, "opt_conv3", true));
    B2_CHECK_CUDA(cudaStreamWaitEvent(sC, ev[2], 0));
    B2_TRY(bwd_op(n, fs, rows, kConv2Wgrad, sC));
    B2_TRY(optimizer_range(n, 1, 1, 1 | 2, rows, sC, "reduce_conv2"));
    B2_TRY(comm_xchg_range(n, 1, 1, 1, sC, "xchg_conv2"));
  }
  B2_TRY(bwd_op(n, fs, rows, kConv2Dgrad, st));
  B2_CHECK_CUDA(cudaEventRecord(ev[3], st));                 // W2 no longer needed
  {
    NoPdlScope side;
    B2_CHECK_CUDA(cudaStreamWaitEvent(sC, ev[3], 0));
    B2_TRY(umma_opt_conv(n, 1, rows, sC, "opt_conv2", true));
  }
  B2_TRY(bwd_op(n, fs, rows, kConv1Wgrad, st));
  {
    NoPdlScope tail;
    B2_TRY(optimizer_range(n, 0, 0, 1 | 2, rows, st, "reduce_conv1"));
    B2_TRY(comm_xchg_range(n, 0, 0, 0, st, "xchg_conv1"));
    B2_TRY(umma_opt_conv(n, 0, rows, st, "opt_conv1", true));
  }
  B2_CHECK_CUDA(cudaEventRecord(ev[4], sA));
  B2_CHECK_CUDA(cudaEventRecord(ev[5], sB));
  B2_CHECK_CUDA(cudaEventRecord(ev[6], sC));
  B2_CHECK_CUDA(cudaStreamWaitEvent(st, ev[4], 0));
  B2_CHECK_CUDA(cudaStreamWaitEvent(st, ev[5], 0));
  B2_CHECK_CUDA(cudaStreamWaitEvent(st, ev[6], 0));
  return B200DQN_OK;
}

// Default data-parallel schedule (comm_p2p.cuh): the step keeps the single-GPU shape.  fc1's gradient is never
// exchanged — every rank gathers all learners' H3 / dZ4 rows (pushed by their producers' successors) and runs
// fc1_wgrad over the global minibatch; the four small layers go through the one-shot LL all-reduce the moment
// their wgrad is done.  Bytes received per step and rank: (W-1) x (0.47 MB planes + 0.64 MB LL lines).
static int backward_and_update_gather(b200dqn_net* n, const FrameSource& fs, int rows, cudaStream_t st) {
  cudaStream_t sA = n->side[0], sB = n->side[1], sC = n->side[2], sN = n->side[3];
  cudaEvent_t* ev = n->ev;
  // experimental: one launch per conv layer for reduce + LL exchange + RMSProp (umma_opt_conv_xll), off by default
  // one launch per conv layer for reduce + LL exchange + update (umma_opt_conv_xll): default since it was validated on
  // hardware at W = 2 (tests/test_gpu_multi.py; ~4.5 us per step); B200DQN_FUSED_XLL=0 restores the three launches
  // Measured: at W = 2 the fused kernel shortens the traced step (87.7 vs 92 us); at W = 8 its 288 polling CTAs per
  // layer spin for ~20 us while the peers catch up and delay the chain's own kernels (profiles/r2n8_timeline_w8.txt),
  // so beyond two ranks the default is the three-launch form with its small polling grid.
  static const int fused_env = getenv("B200DQN_FUSED_XLL") ? atoi(getenv("B200DQN_FUSED_XLL")) : -1;
  const bool fused_xll = fused_env >= 0 ? fused_env != 0 : n->world <= 2;
  B2_CHECK_CUDA(cudaEventRecord(ev[0], st));                 // head done: dZ4 planes, dW5 partials
  B2_CHECK_CUDA(cudaStreamWaitEvent(sA, ev[0], 0));
  {
    NoPdlScope side;
    const bool head_pushed = comm_head_push(n, st, nullptr);  // (opt-in) the head kernel already sent this rank's dZ4 rows
    const bool dz_ll = !head_pushed && comm_dz4_ll_enabled();
    B2_CHECK_CUDA(cudaStreamWaitEvent(sA, ev[14], 0));       // own H3 push (forward, umma_push_h3) has been issued
    if (dz_ll) {
      // all ranks' dZ4 rows in the LL protocol (no flag word, no system fence); the same kernel polls the H3 flags
      B2_TRY(umma_gather_dz4_ll(n, sA));
    } else {
      if (!head_pushed) B2_TRY(umma_push_dz4(n, sA));        // peers' fc1_wgrad wait for these 64 KB
      B2_TRY(comm_wait_pushes(n, sA, head_pushed ? rows : 0));
    }
    B2_TRY(umma_fc1_wgrad_gathered(n, sA));
    // fc2 (8 KB) on the stream the H3 push has left idle: nothing later in the step reads W5
    B2_CHECK_CUDA(cudaStreamWaitEvent(sN, ev[0], 0));
    B2_TRY(cost_finish_on(n, rows, sN));
    B2_TRY(optimizer_range(n, 4, 4, 1 | 2, rows, sN, "reduce_fc2"));
    B2_TRY(comm_xll_layer(n, 4, sN, "xll_fc2"));
    B2_TRY(optimizer_range(n, 4, 4, 4, rows, sN, "opt_fc2"));
    B2_CHECK_CUDA(cudaEventRecord(ev[7], sN));
  }
  B2_TRY(bwd_op(n, fs, rows, kFc1Dgrad, st));
  B2_CHECK_CUDA(cudaEventRecord(ev[1], st));                 // W4 no longer needed
  {
    NoPdlScope side;
    B2_CHECK_CUDA(cudaStreamWaitEvent(sA, ev[1], 0));
    if (n->defer_fc1) B2_TRY(fc1_update_deferred(n, sA));    // applied under the next step's forward (net.cuh)
    else B2_TRY(umma_opt_fc1(n, rows, sA));                  // dW4 is already the global sum
    B2_CHECK_CUDA(cudaStreamWaitEvent(sB, ev[1], 0));
    B2_TRY(bwd_op(n, fs, rows, kConv3Wgrad, sB));
    if (!fused_xll) {
      B2_TRY(optimizer_range(n, 2, 2, 1 | 2, rows, sB, "reduce_conv3"));
      B2_TRY(comm_xll_layer(n, 2, sB, "xll_conv3"));
    }
  }
  B2_TRY(bwd_op(n, fs, rows, kConv3Dgrad, st));
  B2_CHECK_CUDA(cudaEventRecord(ev[2], st));                 // W3 no longer needed
  {
    NoPdlScope side;
    B2_CHECK_CUDA(cudaStreamWaitEvent(sB, ev[2], 0));
    if (fused_xll) B2_TRY(umma_opt_conv_xll(n, 2, rows, sB, "optx_conv3"));
    else B2_TRY(umma_opt_conv(n, 2, rows, sB, "opt_conv3", true));
    B2_CHECK_CUDA(cudaStreamWaitEvent(sC, ev[2], 0));
    B2_TRY(bwd_op(n, fs, rows, kConv2Wgrad, sC));
    if (!fused_xll) {
      B2_TRY(optimizer_range(n, 1, 1, 1 | 2, rows, sC, "reduce_conv2"));
      B2_TRY(comm_xll_layer(n, 1, sC, "xll_conv2"));
    }
  }
  B2_TRY(bwd_op(n, fs, rows, kConv2Dgrad, st));
  B2_CHECK_CUDA(cudaEventRecord(ev[3], st));                 // W2 no longer needed
  {
    NoPdlScope side;
    B2_CHECK_CUDA(cudaStreamWaitEvent(sC, ev[3], 0));
    if (fused_xll) B2_TRY(umma_opt_conv_xll(n, 1, rows, sC, "optx_conv2"));
    else B2_TRY(umma_opt_conv(n, 1, rows, sC, "opt_conv2", true));
  }
  B2_TRY(bwd_op(n, fs, rows, kConv1Wgrad, st));
  {
    NoPdlScope tail;
    if (fused_xll) {
      B2_TRY(umma_opt_conv_xll(n, 0, rows, st, "optx_conv1"));
    } else {
      B2_TRY(optimizer_range(n, 0, 0, 1 | 2, rows, st, "reduce_conv1"));
      B2_TRY(comm_xll_layer(n, 0, st, "xll_conv1"));
      B2_TRY(umma_opt_conv(n, 0, rows, st, "opt_conv1", true));
    }
  }
  B2_CHECK_CUDA(cudaEventRecord(ev[4], sA));
  B2_CHECK_CUDA(cudaEventRecord(ev[5], sB));
  B2_CHECK_CUDA(cudaEventRecord(ev[6], sC));
  B2_CHECK_CUDA(cudaStreamWaitEvent(st, ev[4], 0));
  B2_CHECK_CUDA(cudaStreamWaitEvent(st, ev[5], 0));
  B2_CHECK_CUDA(cudaStreamWaitEvent(st, ev[6], 0));
  B2_CHECK_CUDA(cudaStreamWaitEvent(st, ev[7], 0));
  return B200DQN_OK;
}

static int backward_and_update_multi(b200dqn_net* n, const FrameSource& fs, int rows, cudaStream_t st) {
  if (comm_gather_active(n, st)) return backward_and_update_gather(n, fs, rows, st);
  if (n->xchg_ok && n->xchg_sched == 1) return backward_and_update_xchg(n, fs, rows, st);
  cudaStream_t sA = n->side[0], sB = n->side[1], sC = n->side[2], sN = n->side[3];
  cudaEvent_t* ev = n->ev;
  B2_CHECK_CUDA(cudaEventRecord(ev[0], st));                 // head done: dZ4, dW5 partials
  B2_CHECK_CUDA(cudaStreamWaitEvent(sA, ev[0], 0));
  {
    NoPdlScope side;
    B2_TRY(bwd_op(n, fs, rows, kFc1Wgrad, sA));
    B2_TRY(cost_finish_on(n, rows, sA));
    B2_TRY(optimizer_range(n, 3, 4, 1 | 2, rows, sA, "reduce_fc"));        // partials -> d_g[fc1, fc2]
    B2_CHECK_CUDA(cudaEventRecord(ev[7], sA));
    B2_CHECK_CUDA(cudaStreamWaitEvent(sN, ev[7], 0));
    B2_TRY(comm_allreduce_range(n, 3, 4, sN));
    B2_CHECK_CUDA(cudaEventRecord(ev[8], sN));
  }
  B2_TRY(bwd_op(n, fs, rows, kFc1Dgrad, st));
  B2_CHECK_CUDA(cudaEventRecord(ev[1], st));                 // W4 no longer needed
  {
    NoPdlScope side;
    B2_CHECK_CUDA(cudaStreamWaitEvent(sA, ev[8], 0));
    B2_CHECK_CUDA(cudaStreamWaitEvent(sA, ev[1], 0));
    B2_TRY(umma_opt_fc1(n, rows, sA, true));
    B2_TRY(optimizer_range(n, 4, 4, 4, rows, sA, "opt_fc2"));
    B2_CHECK_CUDA(cudaStreamWaitEvent(sB, ev[1], 0));
    B2_TRY(bwd_op(n, fs, rows, kConv3Wgrad, sB));
  }
  B2_TRY(bwd_op(n, fs, rows, kConv3Dgrad, st));
  B2_CHECK_CUDA(cudaEventRecord(ev[2], st));
  {
    NoPdlScope side;
    B2_TRY(optimizer_range(n, 2, 2, 1 | 2, rows, sB, "reduce_conv3"));     // partials -> d_g, early, off the chain
    B2_CHECK_CUDA(cudaStreamWaitEvent(sC, ev[2], 0));
    B2_TRY(bwd_op(n, fs, rows, kConv2Wgrad, sC));
    B2_TRY(optimizer_range(n, 1, 1, 1 | 2, rows, sC, "reduce_conv2"));
  }
  B2_TRY(bwd_op(n, fs, rows, kConv2Dgrad, st));
  B2_TRY(bwd_op(n, fs, rows, kConv1Wgrad, st));
  {
    NoPdlScope tail;   // kernels around the collective use ordinary dependencies
    B2_TRY(optimizer_range(n, 0, 0, 1 | 2, rows, st, "reduce_conv1"));
    B2_CHECK_CUDA(cudaEventRecord(ev[5], sB));
    B2_CHECK_CUDA(cudaEventRecord(ev[6], sC));
    B2_CHECK_CUDA(cudaEventRecord(ev[9], st));
    B2_CHECK_CUDA(cudaStreamWaitEvent(sN, ev[5], 0));
    B2_CHECK_CUDA(cudaStreamWaitEvent(sN, ev[6], 0));
    B2_CHECK_CUDA(cudaStreamWaitEvent(sN, ev[9], 0));
    B2_TRY(comm_allreduce_range(n, 0, 2, sN));                              // one small collective for conv1..3
    B2_CHECK_CUDA(cudaEventRecord(ev[10], sN));
    B2_CHECK_CUDA(cudaStreamWaitEvent(st, ev[10], 0));
    B2_CHECK_CUDA(cudaStreamWaitEvent(sB, ev[10], 0));
    B2_CHECK_CUDA(cudaStreamWaitEvent(sC, ev[10], 0));
    B2_TRY(umma_opt_conv(n, 2, rows, sB, "opt_conv3", true));               // the three tiny updates run side by side
    B2_TRY(umma_opt_conv(n, 1, rows, sC, "opt_conv2", true));
    B2_TRY(umma_opt_conv(n, 0, rows, st, "opt_conv1", true));
    B2_CHECK_CUDA(cudaEventRecord(ev[11], sB));
    B2_CHECK_CUDA(cudaEventRecord(ev[12], sC));
    B2_CHECK_CUDA(cudaStreamWaitEvent(st, ev[11], 0));
    B2_CHECK_CUDA(cudaStreamWaitEvent(st, ev[12], 0));
  }
  B2_CHECK_CUDA(cudaEventRecord(ev[4], sA));
  B2_CHECK_CUDA(cudaStreamWaitEvent(st, ev[4], 0));
  return B200DQN_OK;
}

static int backward_and_update(b200dqn_net* n, const FrameSource& fs, int rows, cudaStream_t st, bool update) {
  if (update && n->world > 1 && !g_prof_on && n->use_branches && st != nullptr &&
      n->cfg.math_mode == B200DQN_MATH_TCGEN05)
    return backward_and_update_multi(n, fs, rows, st);
  // Under the event profiler the same kernels run, but every "branch" is the main stream (serialised).
  const bool branches = update && n->world == 1 && n->use_branches && (st != nullptr || g_prof_on);
  if (!branches) {
    for (int op = kFc1Wgrad; op <= kConv1Wgrad; ++op) B2_TRY(bwd_op(n, fs, rows, BwdOp(op), st));
    B2_TRY(cost_finish_on(n, rows, st));
    if (!update) return B200DQN_OK;
    if (n->world > 1) {
      B2_TRY(optimizer_range(n, 0, kLayers - 1, 1 | 2, rows, st, "grad_reduce"));
      B2_TRY(comm_allreduce_grads(n, st));
      B2_PROF("allreduce", st);
      B2_TRY(optimizer_range(n, 0, kLayers - 1, 4, rows, st, "optimizer"));
    } else {
      B2_TRY(optimizer_range(n, 0, kLayers - 1, 1 | 4, rows, st, "optimizer"));
    }
    return B200DQN_OK;
  }
  cudaStream_t sA = g_prof_on ? st : n->side[0], sB = g_prof_on ? st : n->side[1], sC = g_prof_on ? st : n->side[2];
  cudaStream_t sN = g_prof_on ? st : n->side[3];
  cudaEvent_t* ev = n->ev;
  const bool tc = n->cfg.math_mode == B200DQN_MATH_TCGEN05;
  static const bool fc1_fused_epilogue = getenv("B200DQN_FC1_FUSED") != nullptr;   // experimental alternative
  // experiment knobs: bit op of B200DQN_NOPDL_OPS launches that chain kernel without the programmatic dependency;
  // B200DQN_WGRAD_ONE_STREAM puts conv2_wgrad / opt_conv2 on conv3_wgrad's branch
  static const int nopdl_ops = getenv("B200DQN_NOPDL_OPS") ? int(strtol(getenv("B200DQN_NOPDL_OPS"), nullptr, 0)) : 0;
  if (getenv("B200DQN_WGRAD_ONE_STREAM") && !g_prof_on) sC = sB;
  auto chain_op = [&](BwdOp op) -> int {
    if (nopdl_ops >> int(op) & 1) {
      NoPdlScope plain;
      return bwd_op(n, fs, rows, op, st);
    }
    return bwd_op(n, fs, rows, op, st);
  };
  B2_CHECK_CUDA(cudaEventRecord(ev[0], st));                 // dZ4, the dW5 partials and the per-sample costs are ready
  B2_CHECK_CUDA(cudaStreamWaitEvent(sA, ev[0], 0));
  B2_CHECK_CUDA(cudaStreamWaitEvent(sN, ev[0], 0));
  {
    NoPdlScope side;
    if (!(tc && fc1_fused_epilogue)) B2_TRY(bwd_op(n, fs, rows, kFc1Wgrad, sA));   // overlaps fc1_dgrad (few CTAs)
    // fourth branch: the scalar cost and the 512 x A layer — nothing later in the step reads W5, and nothing here
    // sits in front of the fc1 optimizer any more
    B2_TRY(cost_finish_on(n, rows, sN));
    if (tc) B2_TRY(opt_fc2_small(n, rows, sN));
  }
  B2_TRY(chain_op(kFc1Dgrad));
  B2_CHECK_CUDA(cudaEventRecord(ev[1], st));                 // dZ3 ready, W4 no longer needed
  // experiment knob B200DQN_OPT_FC1_WHEN: "ev2" / "ev3" = also wait for conv3_dgrad / conv2_dgrad, "last" = same
  // dependencies but captured after every other node of the step, "skip" = not launched (timing studies only)
  static const char* fc1_when_env = getenv("B200DQN_OPT_FC1_WHEN");
  const int fc1_when = !fc1_when_env || g_prof_on ? 0 : !strcmp(fc1_when_env, "ev2") ? 2 : !strcmp(fc1_when_env, "ev3") ? 3
                       : !strcmp(fc1_when_env, "last") ? 4 : !strcmp(fc1_when_env, "skip") ? 5 : 0;
  auto opt_fc1_now = [&]() -> int {
    NoPdlScope side;
    if (tc && fc1_fused_epilogue) return umma_fc1_wgrad_fused(n, rows, sA, n->keep_grads);
    if (tc && n->defer_fc1) return fc1_update_deferred(n, sA);   // applied under the next step's forward (net.cuh)
    if (tc) return umma_opt_fc1(n, rows, sA);                // smem-free: co-resides with the dgrad chain
    return optimizer_range(n, 3, 4, 1 | 4, rows, sA, "opt_fc");
  };
  B2_CHECK_CUDA(cudaStreamWaitEvent(sA, ev[1], 0));
  if (fc1_when == 0) B2_TRY(opt_fc1_now());
  B2_CHECK_CUDA(cudaStreamWaitEvent(sB, ev[1], 0));
  { NoPdlScope side; B2_TRY(bwd_op(n, fs, rows, kConv3Wgrad, sB)); }
  B2_TRY(chain_op(kConv3Dgrad));
  B2_CHECK_CUDA(cudaEventRecord(ev[2], st));                 // dZ2 ready, W3 no longer needed
  if (fc1_when == 2) {
    B2_CHECK_CUDA(cudaStreamWaitEvent(sA, ev[2], 0));
    B2_TRY(opt_fc1_now());
  }
  B2_CHECK_CUDA(cudaStreamWaitEvent(sB, ev[2], 0));
  {
    NoPdlScope side;
    if (n->cfg.math_mode == B200DQN_MATH_TCGEN05) B2_TRY(umma_opt_conv(n, 2, rows, sB, "opt_conv3"));
    else B2_TRY(optimizer_range(n, 2, 2, 1 | 4, rows, sB, "opt_conv3"));
  }
  B2_CHECK_CUDA(cudaStreamWaitEvent(sC, ev[2], 0));
  { NoPdlScope side; B2_TRY(bwd_op(n, fs, rows, kConv2Wgrad, sC)); }
  B2_TRY(chain_op(kConv2Dgrad));
  B2_CHECK_CUDA(cudaEventRecord(ev[3], st));                 // dZ1 ready, W2 no longer needed
  if (fc1_when == 3) {
    B2_CHECK_CUDA(cudaStreamWaitEvent(sA, ev[3], 0));
    B2_TRY(opt_fc1_now());
  }
  B2_CHECK_CUDA(cudaStreamWaitEvent(sC, ev[3], 0));
  {
    NoPdlScope side;
    if (n->cfg.math_mode == B200DQN_MATH_TCGEN05) B2_TRY(umma_opt_conv(n, 1, rows, sC, "opt_conv2"));
    else B2_TRY(optimizer_range(n, 1, 1, 1 | 4, rows, sC, "opt_conv2"));
  }
  B2_TRY(chain_op(kConv1Wgrad));
  if (n->cfg.math_mode == B200DQN_MATH_TCGEN05) B2_TRY(umma_opt_conv(n, 0, rows, st, "opt_conv1"));
  else B2_TRY(optimizer_range(n, 0, 0, 1 | 4, rows, st, "opt_conv1"));
  if (fc1_when == 4) B2_TRY(opt_fc1_now());
  B2_CHECK_CUDA(cudaEventRecord(ev[4], sA));
  B2_CHECK_CUDA(cudaEventRecord(ev[5], sB));
  B2_CHECK_CUDA(cudaEventRecord(ev[6], sC));
  B2_CHECK_CUDA(cudaEventRecord(ev[7], sN));
  B2_CHECK_CUDA(cudaStreamWaitEvent(st, ev[4], 0));
  B2_CHECK_CUDA(cudaStreamWaitEvent(st, ev[5], 0));
  B2_CHECK_CUDA(cudaStreamWaitEvent(st, ev[6], 0));
  B2_CHECK_CUDA(cudaStreamWaitEvent(st, ev[7], 0));
  return B200DQN_OK;
}

// One DeepQNetwork.train on device-resident inputs (the caller counts train_iterations, :168).
static int train_step(b200dqn_net* n, const FrameSource& fs, const uint8_t* actions, const int64_t* rewards,
                      const uint8_t* terminals, const int32_t* midx, cudaStream_t st) {
  const int rows = n->nb;
  HeadTrainArgs td{1, actions, rewards, terminals, midx, n->cfg.discount_rate, n->cfg.min_reward, n->cfg.max_reward,
                   float(n->cfg.clip_error), n->d_delta, n->d_step, n->d_rowcost, n->d_dz4,
                   n->d_part + n->lt.part_off[4], nullptr, 0,
                   n->cfg.optimizer == B200DQN_OPT_ADAM ? n->d_optscal : nullptr, float(n->cfg.learning_rate), n->A,
                   reinterpret_cast<uint32_t*>(n->d_cost + kCostRing + 1), HeadPush{}};
  umma_dz4_planes(n, &td.dz4_hi, &td.dz4_lo_off);
  comm_head_push(n, st, &td.push);
  B2_TRY(forward(n, fs, 2, rows, st, td));
  return backward_and_update(n, fs, rows, st, true);
}

// ---------------------------------------------------------------- layout conversion (host)
// Neon layout <-> internal layout index map for one layer; returns internal linear index.
static inline int64_t neon_to_internal(int layer, int64_t i, int A) {
  switch (layer) {
    case 0: return i;  // (c,r,s) x K: identical
    case 1: {          // neon rows (c,r,s), C=32,R=4 -> internal rows (r,s,c)
      const int64_t k = i % kC2, row = i / kC2;
      const int c = int(row / 16), r = int(row / 4) % 4, s = int(row % 4);
      return ((int64_t(r) * 4 + s) * kC1 + c) * kC2 + k;
    }
    case 2: {          // C=64, R=3
      const int64_t k = i % kC3, row = i / kC3;
      const int c = int(row / 9), r = int(row / 3) % 3, s = int(row % 3);
      return ((int64_t(r) * 3 + s) * kC2 + c) * kC3 + k;
    }
    case 3: {          // neon W[n][(c,p,q)] -> internal W[(p,q,c)][n]
      const int64_t nn = i / kFlat, col = i % kFlat;
      const int c = int(col / 49), p = int(col / 7) % 7, q = int(col % 7);
      return ((int64_t(p) * 7 + q) * kC3 + c) * kHidden + nn;
    }
    default: {         // neon W[a][k] -> internal W[k][a]
      const int64_t a = i / kHidden, k = i % kHidden;
      return k * A + a;
    }
  }
}

}  // namespace b200

using namespace b200;

// ============================================================================ C ABI: network
extern "C" int b200dqn_net_config_default(b200dqn_net_config* cfg, int num_actions) {
  B2_REQUIRE(cfg, B200DQN_EINVAL, "null cfg");
  memset(cfg, 0, sizeof(*cfg));
  cfg->num_actions = num_actions;
  cfg->batch_size = 32;          // main.py:39
  cfg->history_length = 4;       // main.py:34
  cfg->screen_h = cfg->screen_w = 84;  // main.py:27-28
  cfg->discount_rate = 0.99;     // main.py:38
  cfg->learning_rate = 0.00025;  // main.py:37
  cfg->decay_rate = 0.95;        // main.py:41
  cfg->clip_error = 1.0;         // main.py:42
  cfg->min_reward = -1;          // main.py:43
  cfg->max_reward = 1;           // main.py:44
  cfg->target_steps = 10000;     // main.py:63
  cfg->math_mode = B200DQN_MATH_FP32_SIMT;
  cfg->optimizer = B200DQN_OPT_RMSPROP;   // main.py:40
  return B200DQN_OK;
}

extern "C" int b200dqn_net_create(int device, const b200dqn_net_config* cfg, b200dqn_net** out) {
  B2_REQUIRE(cfg && out, B200DQN_EINVAL, "net_create: null argument");
  B2_REQUIRE(cfg->num_actions >= 1 && cfg->num_actions <= kMaxActions, B200DQN_EINVAL,
             "net_create: num_actions %d not in [1,%d]", cfg->num_actions, kMaxActions);
  B2_REQUIRE(cfg->batch_size >= 1 && cfg->batch_size <= 4096, B200DQN_EINVAL, "net_create: batch_size");
  B2_REQUIRE(cfg->screen_h == kFrameH && cfg->screen_w == kFrameW && cfg->history_length == kHist,
             B200DQN_ENOTIMPL,
             "net_create: only the reference's 84x84x4 Nature-DQN geometry is implemented (got %dx%dx%d)",
             cfg->screen_h, cfg->screen_w, cfg->history_length);
  B2_REQUIRE(cfg->math_mode == B200DQN_MATH_FP32_SIMT || cfg->math_mode == B200DQN_MATH_TCGEN05, B200DQN_EINVAL,
             "net_create: unknown math_mode %d", cfg->math_mode);
  B2_REQUIRE(cfg->optimizer >= B200DQN_OPT_RMSPROP && cfg->optimizer <= B200DQN_OPT_ADADELTA, B200DQN_EINVAL,
             "net_create: unknown optimizer %d", cfg->optimizer);   // deepqnetwork.py:61 `assert false, "Unknown optimizer"`
  DeviceGuard g(device);
  auto* n = new (std::nothrow) b200dqn_net();
  B2_REQUIRE(n, B200DQN_EINVAL, "out of host memory");
  n->n_states = cfg->optimizer == B200DQN_OPT_ADAM ? 2 : cfg->optimizer == B200DQN_OPT_ADADELTA ? 3 : 1;
  n->device = device;
  B2_CHECK_CUDA(cudaDeviceGetAttribute(&n->sm_count, cudaDevAttrMultiProcessorCount, device));
  n->cfg = *cfg;
  n->nb = cfg->batch_size;
  n->A = cfg->num_actions;
  const int nb = n->nb, A = n->A;
  LayerTable& lt = n->lt;
  const int rows_[kLayers] = {kK1, kK2, kK3, kFlat, kHidden};
  const int cols_[kLayers] = {kC1, kC2, kC3, kHidden, A};
  lt.off[0] = 0;
  for (int l = 0; l < kLayers; ++l) {
    lt.rows[l] = rows_[l];
    lt.cols[l] = cols_[l];
    lt.off[l + 1] = lt.off[l] + int64_t(rows_[l]) * cols_[l];
  }
  n->n_params = lt.off[kLayers];
  B2_REQUIRE(n->n_params % 4 == 0, B200DQN_EINVAL, "parameter count must be a multiple of 4");
  const int kred[3] = {nb * kP1 * kP1, nb * kP2 * kP2, nb * kP3 * kP3};
  const int base[3] = {512, 96, 112};
  int64_t po = 0;
  for (int l = 0; l < kLayers; ++l) {
    if (l == 4)
      lt.splits[l] = nb;   // the head kernel leaves one dW5 partial per sample
    else if (cfg->math_mode == B200DQN_MATH_TCGEN05)
      lt.splits[l] = l < 3 ? umma_wgrad_splits(l, nb) : 1;
    else
      lt.splits[l] = l < 3 ? int(cdiv(kred[l], wgrad_chunk(kred[l], base[l]))) : 1;
    lt.part_off[l] = po;
    po += int64_t(lt.splits[l]) * (lt.off[l + 1] - lt.off[l]);
  }
  n->part_elems = po;

  auto fmalloc = [&](float** p, size_t elems) -> cudaError_t {
    cudaError_t e = cudaMalloc(p, elems * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(*p, 0, elems * sizeof(float));
    return e;
  };
  B2_CHECK_CUDA(fmalloc(&n->d_w, n->n_params));
  B2_CHECK_CUDA(fmalloc(&n->d_s, n->n_params * n->n_states));
  B2_CHECK_CUDA(fmalloc(&n->d_optscal, 4));
  if (cfg->target_steps) {
    B2_CHECK_CUDA(fmalloc(&n->d_tw, n->n_params));
    B2_CHECK_CUDA(fmalloc(&n->d_ts, n->n_params * n->n_states));
  } else {
    n->d_tw = n->d_w;  // deepqnetwork.py:72-73: the target model IS the online model
    n->d_ts = n->d_s;
  }
  // the gradient buffer is the one allocation peers map (comm.cu): their flag words sit behind it
  B2_CHECK_CUDA(fmalloc(&n->d_g, n->n_params + kXFlagWords));
  n->d_xflags = reinterpret_cast<uint32_t*>(n->d_g + n->n_params);
  constexpr int kXWords = kXChannels * kXMaxBlocks + 1 + 2 * kXChannels + 2 * kXPushChannels + 1;
  B2_CHECK_CUDA(cudaMalloc(&n->d_xepoch, kXWords * sizeof(uint32_t)));
  B2_CHECK_CUDA(cudaMemset(n->d_xepoch, 0, kXWords * sizeof(uint32_t)));
  n->d_xerr = n->d_xepoch + kXChannels * kXMaxBlocks;
  n->d_xll_epoch = n->d_xerr + 1;
  n->d_xpush_epoch = n->d_xll_epoch + 2 * kXChannels;
  B2_CHECK_CUDA(fmalloc(&n->d_part, n->part_elems));
  for (int z = 0; z < 2; ++z) {
    B2_CHECK_CUDA(fmalloc(&n->d_h1[z], size_t(nb) * kP1 * kP1 * kC1));
    B2_CHECK_CUDA(fmalloc(&n->d_h2[z], size_t(nb) * kP2 * kP2 * kC2));
    B2_CHECK_CUDA(fmalloc(&n->d_h3[z], size_t(nb) * kFlat));
    B2_CHECK_CUDA(fmalloc(&n->d_h4[z], size_t(nb) * kHidden));
    B2_CHECK_CUDA(fmalloc(&n->d_q[z], size_t(nb) * A));
  }
  B2_CHECK_CUDA(fmalloc(&n->d_fc1part, size_t(2) * kFc1Splits * nb * kHidden));
  B2_CHECK_CUDA(fmalloc(&n->d_delta, size_t(nb) * A));
  B2_CHECK_CUDA(fmalloc(&n->d_dz4, size_t(nb) * kHidden));
  B2_CHECK_CUDA(fmalloc(&n->d_dz3, size_t(nb) * kFlat));
  B2_CHECK_CUDA(fmalloc(&n->d_dz2, size_t(nb) * kP2 * kP2 * kC2));
  B2_CHECK_CUDA(fmalloc(&n->d_dz1, size_t(nb) * kP1 * kP1 * kC1));
  B2_CHECK_CUDA(fmalloc(&n->d_cost, kCostRing + 2));   // ring, "latest" slot, action-range flag word
  B2_CHECK_CUDA(cudaMalloc(&n->d_step, sizeof(uint32_t)));
  B2_CHECK_CUDA(cudaMemset(n->d_step, 0, sizeof(uint32_t)));
  B2_CHECK_CUDA(fmalloc(&n->d_rowcost, nb));
  const size_t state_bytes = size_t(nb) * kHist * kFrameBytes;
  B2_CHECK_CUDA(cudaMalloc(&n->d_pre, state_bytes + 256));
  B2_CHECK_CUDA(cudaMalloc(&n->d_post, state_bytes + 256));
  B2_CHECK_CUDA(cudaMalloc(&n->d_act, nb));
  B2_CHECK_CUDA(cudaMalloc(&n->d_term, nb));
  B2_CHECK_CUDA(cudaMalloc(&n->d_rew, nb * sizeof(int64_t)));
  B2_CHECK_CUDA(cudaMalloc(&n->d_iota1, nb * sizeof(int32_t)));
  B2_CHECK_CUDA(cudaMalloc(&n->d_iota4, nb * sizeof(int32_t)));
  k_iota<<<cdiv(nb, 128), 128>>>(n->d_iota1, n->d_iota4, nb, kHist);
  B2_LAUNCH_CHECK();
  n->pin_bytes = 2 * state_bytes + size_t(nb) * 16 + size_t(nb) * A * sizeof(float) + 256;
  B2_CHECK_CUDA(cudaMallocHost(&n->h_pin, n->pin_bytes));
  {
    void* m = nullptr;
    B2_CHECK_CUDA(cudaHostAlloc(&m, 4096, cudaHostAllocMapped));
    memset(m, 0, 4096);
    n->h_res = static_cast<volatile uint32_t*>(m);
  }
  {
    int prio_lo = 0, prio_hi = 0;   // numerically larger = lower priority
    B2_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    const char* sp = getenv("B200DQN_SIDE_PRIO");            // experiment knob: "hi" / "lo" for all four
    for (int i = 0; i < 4; ++i) {
      int prio = i < 3 ? prio_lo : prio_hi;
      if (sp && !strcmp(sp, "hi")) prio = prio_hi;
      if (sp && !strcmp(sp, "lo")) prio = prio_lo;
      B2_CHECK_CUDA(cudaStreamCreateWithPriority(&n->side[i], cudaStreamNonBlocking, prio));
    }
  }
  for (auto& e : n->ev) B2_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  B2_CHECK_CUDA(cudaMalloc(&n->d_fc1_pending, sizeof(uint32_t)));
  B2_CHECK_CUDA(cudaMemset(n->d_fc1_pending, 0, sizeof(uint32_t)));
  n->use_graph = getenv("B200DQN_NO_GRAPH") == nullptr;
  n->use_branches = getenv("B200DQN_NO_BRANCHES") == nullptr;
  int rc = umma_net_init(n);
  if (rc) return rc;
  B2_CHECK_CUDA(cudaDeviceSynchronize());
  *out = n;
  return B200DQN_OK;
}

extern "C" int b200dqn_net_destroy(b200dqn_net* n) {
  if (!n) return B200DQN_OK;
  DeviceGuard g(n->device);
  cudaDeviceSynchronize();
  comm_destroy(n);
  umma_net_destroy(n);
  destroy_step_graphs(n);
  if (n->graph_train_exec) cudaGraphExecDestroy(n->graph_train_exec);
  cudaFree(n->d_fc1_pending);
  if (n->graph_predict_exec) cudaGraphExecDestroy(n->graph_predict_exec);
  for (auto& sd : n->side) if (sd) cudaStreamDestroy(sd);
  for (auto& e : n->ev) if (e) cudaEventDestroy(e);
  if (n->d_tw != n->d_w) { cudaFree(n->d_tw); cudaFree(n->d_ts); }
  cudaFree(n->d_optscal);
  cudaFree(n->d_w); cudaFree(n->d_s); cudaFree(n->d_g); cudaFree(n->d_part); cudaFree(n->d_xepoch);
  for (int z = 0; z < 2; ++z) {
    cudaFree(n->d_h1[z]); cudaFree(n->d_h2[z]); cudaFree(n->d_h3[z]); cudaFree(n->d_h4[z]); cudaFree(n->d_q[z]);
  }
  cudaFree(n->d_fc1part); cudaFree(n->d_delta); cudaFree(n->d_dz4); cudaFree(n->d_dz3); cudaFree(n->d_dz2);
  cudaFree(n->d_dz1); cudaFree(n->d_cost); cudaFree(n->d_step); cudaFree(n->d_rowcost); cudaFree(n->d_pre); cudaFree(n->d_post);
  cudaFree(n->d_act); cudaFree(n->d_term); cudaFree(n->d_rew); cudaFree(n->d_iota1); cudaFree(n->d_iota4);
  cudaFreeHost(n->h_pin);
  cudaFreeHost(const_cast<uint32_t*>(n->h_res));
  delete n;
  return B200DQN_OK;
}

extern "C" int b200dqn_net_layer_shape(const b200dqn_net* n, int layer, int* rows, int* cols) {
  B2_REQUIRE(n && layer >= 0 && layer < kLayers, B200DQN_EINVAL, "net_layer_shape: bad layer");
  // NEON shapes: conv (C*R*S, K); linear (nout, nin)
  const int r[kLayers] = {kK1, kK2, kK3, kHidden, n->A};
  const int c[kLayers] = {kC1, kC2, kC3, kFlat, kHidden};
  if (rows) *rows = r[layer];
  if (cols) *cols = c[layer];
  return B200DQN_OK;
}

static int xfer_params(b200dqn_net* n, float* dev_base, int layer, float* host, bool to_device, cudaStream_t st) {
  const int64_t off = n->lt.off[layer], cnt = n->lt.off[layer + 1] - off;
  std::vector<float> tmp(cnt);
  if (to_device) {
    for (int64_t i = 0; i < cnt; ++i) tmp[neon_to_internal(layer, i, n->A)] = host[i];
    B2_CHECK_CUDA(cudaMemcpyAsync(dev_base + off, tmp.data(), cnt * sizeof(float), cudaMemcpyHostToDevice, st));
    B2_CHECK_CUDA(cudaStreamSynchronize(st));
  } else {
    B2_CHECK_CUDA(cudaMemcpyAsync(tmp.data(), dev_base + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, st));
    B2_CHECK_CUDA(cudaStreamSynchronize(st));
    for (int64_t i = 0; i < cnt; ++i) host[i] = tmp[neon_to_internal(layer, i, n->A)];
  }
  return B200DQN_OK;
}

extern "C" int b200dqn_net_set_weights(b200dqn_net* n, int which, int layer, const float* host_W,
                                       const float* host_S, void* stream) {
  B2_REQUIRE(n && host_W && layer >= 0 && layer < kLayers && (which == 0 || which == 1), B200DQN_EINVAL,
             "net_set_weights: bad argument");
  DeviceGuard g(n->device);
  cudaStream_t st = as_stream(stream);
  int rc = xfer_params(n, which ? n->d_tw : n->d_w, layer, const_cast<float*>(host_W), true, st);
  if (rc) return rc;
  if (host_S && (rc = xfer_params(n, which ? n->d_ts : n->d_s, layer, const_cast<float*>(host_S), true, st))) return rc;
  return umma_weights_changed(n, st);
}

extern "C" int b200dqn_net_get_weights(b200dqn_net* n, int which, int layer, float* host_W, float* host_S,
                                       void* stream) {
  B2_REQUIRE(n && layer >= 0 && layer < kLayers && (which == 0 || which == 1), B200DQN_EINVAL,
             "net_get_weights: bad argument");
  DeviceGuard g(n->device);
  cudaStream_t st = as_stream(stream);
  int rc;
  if (host_W && (rc = xfer_params(n, which ? n->d_tw : n->d_w, layer, host_W, false, st))) return rc;
  if (host_S && (rc = xfer_params(n, which ? n->d_ts : n->d_s, layer, host_S, false, st))) return rc;
  return B200DQN_OK;
}

extern "C" int b200dqn_net_num_states(const b200dqn_net* n, int* count) {
  B2_REQUIRE(n && count, B200DQN_EINVAL, "net_num_states: null argument");
  *count = n->n_states;
  return B200DQN_OK;
}

extern "C" int b200dqn_net_set_state(b200dqn_net* n, int which, int layer, int k, const float* host_S, void* stream) {
  B2_REQUIRE(n && host_S && layer >= 0 && layer < kLayers && (which == 0 || which == 1) && k >= 0 && k < n->n_states,
             B200DQN_EINVAL, "net_set_state: bad argument");
  DeviceGuard g(n->device);
  return xfer_params(n, (which ? n->d_ts : n->d_s) + int64_t(k) * n->n_params, layer, const_cast<float*>(host_S), true,
                     as_stream(stream));
}

extern "C" int b200dqn_net_get_state(b200dqn_net* n, int which, int layer, int k, float* host_S, void* stream) {
  B2_REQUIRE(n && host_S && layer >= 0 && layer < kLayers && (which == 0 || which == 1) && k >= 0 && k < n->n_states,
             B200DQN_EINVAL, "net_get_state: bad argument");
  DeviceGuard g(n->device);
  return xfer_params(n, (which ? n->d_ts : n->d_s) + int64_t(k) * n->n_params, layer, host_S, false, as_stream(stream));
}

extern "C" int b200dqn_net_sync_target(b200dqn_net* n, void* stream) {
  B2_REQUIRE(n, B200DQN_EINVAL, "null net");
  if (n->d_tw == n->d_w) return B200DQN_OK;  // target_steps == 0: alias
  DeviceGuard g(n->device);
  cudaStream_t st = as_stream(stream);
  B2_CHECK_CUDA(cudaMemcpyAsync(n->d_tw, n->d_w, n->n_params * sizeof(float), cudaMemcpyDeviceToDevice, st));
  B2_CHECK_CUDA(cudaMemcpyAsync(n->d_ts, n->d_s, n->n_params * n->n_states * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return umma_target_synced(n, st);
}

extern "C" int b200dqn_net_predict_device(b200dqn_net* n, const uint8_t* dev_states, int live_rows, float* dev_q,
                                          void* stream) {
  B2_REQUIRE(n && dev_states && dev_q && live_rows >= 1 && live_rows <= n->nb, B200DQN_EINVAL,
             "net_predict_device: bad argument");
  DeviceGuard g(n->device);
  cudaStream_t st = as_stream(stream);
  const int64_t state_frames = int64_t(n->nb) * kHist;
  FrameSource fs{{dev_states, dev_states}, {n->d_iota4, n->d_iota4}, {0, 0}, {state_frames, state_frames}};
  HeadTrainArgs no_td{};
  int rc = forward(n, fs, 1, live_rows, st, no_td);
  if (rc) return rc;
  if (dev_q != n->d_q[0])
    B2_CHECK_CUDA(cudaMemcpyAsync(dev_q, n->d_q[0], size_t(live_rows) * n->A * sizeof(float),
                                  cudaMemcpyDeviceToDevice, st));
  if (live_rows < n->nb) {
    prefer_max_smem(k_zero_rows);
    k_zero_rows<<<cdiv((n->nb - live_rows) * n->A, 128), 128, 0, st>>>(dev_q, live_rows, n->nb, n->A);
    B2_LAUNCH_CHECK();
  }
  return B200DQN_OK;
}

// Q rows of a fast-path predict -> host-mapped memory: data, system fence, sequence number
__global__ void k_publish_q_counter(const float* __restrict__ q, int count, volatile uint32_t* host_res, uint32_t* counter) {
  for (int i = threadIdx.x; i < count; i += blockDim.x) host_res[kHostQ + i] = __float_as_uint(q[i]);
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t seq = *counter + 1;     // device-resident count of fast-path predicts: the graph replays unchanged
    *counter = seq;
    __threadfence_system();
    host_res[2] = seq;
  }
}

// agent.py:55-61 on a device-resident state window (StateBuffer): the forward pass for the live rows, captured once
// into a CUDA graph (one launch per env step), Q rows back through host-mapped memory (no memcpy, polled).
// host_q receives (batch, A); rows >= live_rows are exact zeros (no biases: Q(0) = 0).
extern "C" int b200dqn_net_predict_device_host(b200dqn_net* n, const uint8_t* dev_states, int live_rows, float* host_q,
                                               void* stream) {
  B2_REQUIRE(n && dev_states && host_q && live_rows >= 1 && live_rows <= n->nb, B200DQN_EINVAL,
             "net_predict_device_host: bad argument");
  DeviceGuard g(n->device);
  cudaStream_t st = as_stream(stream);
  const int count = live_rows * n->A;
  if (count > kHostQFloats || st == nullptr || !n->use_graph || g_prof_on) {   // general path: device predict + copy
    int rc = b200dqn_net_predict_device(n, dev_states, live_rows, n->d_q[0], stream);
    if (rc) return rc;
    B2_CHECK_CUDA(cudaMemcpyAsync(host_q, n->d_q[0], size_t(n->nb) * n->A * sizeof(float), cudaMemcpyDeviceToHost, st));
    B2_CHECK_CUDA(cudaStreamSynchronize(st));
    return B200DQN_OK;
  }
  if (!n->graph_predict_exec || n->graph_predict_states != dev_states || n->graph_predict_rows != live_rows ||
      n->graph_predict_stream != st) {
    if (n->graph_predict_exec) { cudaGraphExecDestroy(n->graph_predict_exec); n->graph_predict_exec = nullptr; }
    cudaGraph_t graph = nullptr;
    B2_CHECK_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int64_t state_frames = int64_t(n->nb) * kHist;
    FrameSource fs{{dev_states, dev_states}, {n->d_iota4, n->d_iota4}, {0, 0}, {state_frames, state_frames}};
    HeadTrainArgs no_td{};
    int rc = forward(n, fs, 1, live_rows, st, no_td);
    if (!rc) {
      prefer_max_smem(k_publish_q_counter);
      k_publish_q_counter<<<1, 64, 0, st>>>(n->d_q[0], count, n->h_res, n->d_optscal_u32());
      if (cudaGetLastError() != cudaSuccess) rc = B200DQN_ECUDA;
    }
    cudaError_t e = cudaStreamEndCapture(st, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    B2_CHECK_CUDA(e);
    B2_CHECK_CUDA(cudaGraphInstantiate(&n->graph_predict_exec, graph, 0));
    cudaGraphDestroy(graph);
    n->graph_predict_states = dev_states; n->graph_predict_rows = live_rows; n->graph_predict_stream = st;
  }
  B2_CHECK_CUDA(cudaGraphLaunch(n->graph_predict_exec, st));
  n->predicts_launched += 1;
  int rc = poll_mapped_seq(n->h_res + 2, n->predicts_launched, st, "predict result");
  if (rc) return rc;
  for (int i = 0; i < count; ++i) {
    const uint32_t bits = n->h_res[kHostQ + i];
    memcpy(host_q + i, &bits, sizeof(float));
  }
  memset(host_q + count, 0, (size_t(n->nb) * n->A - count) * sizeof(float));
  return B200DQN_OK;
}

extern "C" int b200dqn_net_predict(b200dqn_net* n, const uint8_t* host_states, float* host_q, void* stream) {
  B2_REQUIRE(n && host_states && host_q, B200DQN_EINVAL, "net_predict: null argument");
  DeviceGuard g(n->device);
  cudaStream_t st = as_stream(stream);
  const size_t state_bytes = size_t(n->nb) * kHist * kFrameBytes;
  memcpy(n->h_pin, host_states, state_bytes);
  B2_CHECK_CUDA(cudaMemcpyAsync(n->d_pre, n->h_pin, state_bytes, cudaMemcpyHostToDevice, st));
  int rc = b200dqn_net_predict_device(n, n->d_pre, n->nb, n->d_q[0], stream);
  if (rc) return rc;
  float* hq = reinterpret_cast<float*>(n->h_pin + 2 * state_bytes + size_t(n->nb) * 16);
  B2_CHECK_CUDA(cudaMemcpyAsync(hq, n->d_q[0], size_t(n->nb) * n->A * sizeof(float), cudaMemcpyDeviceToHost, st));
  B2_CHECK_CUDA(cudaStreamSynchronize(st));
  memcpy(host_q, hq, size_t(n->nb) * n->A * sizeof(float));
  return B200DQN_OK;
}

extern "C" int b200dqn_net_train_device(b200dqn_net* n, const uint8_t* dev_pre, const uint8_t* dev_actions,
                                        const int64_t* dev_rewards, const uint8_t* dev_post,
                                        const uint8_t* dev_terminals, void* stream) {
  B2_REQUIRE(n && dev_pre && dev_actions && dev_rewards && dev_post && dev_terminals, B200DQN_EINVAL,
             "net_train_device: null argument");
  DeviceGuard g(n->device);
  const int64_t state_frames = int64_t(n->nb) * kHist;
  FrameSource fs{{dev_pre, dev_post}, {n->d_iota4, n->d_iota4}, {0, 0}, {state_frames, state_frames}};
  B2_TRY(train_step(n, fs, dev_actions, dev_rewards, dev_terminals, n->d_iota1, as_stream(stream)));
  n->train_iterations += 1;
  return B200DQN_OK;
}

extern "C" int b200dqn_net_train(b200dqn_net* n, const uint8_t* host_pre, const uint8_t* host_actions,
                                 const int64_t* host_rewards, const uint8_t* host_post, const uint8_t* host_terminals,
                                 float* host_cost, void* stream) {
  B2_REQUIRE(n && host_pre && host_actions && host_rewards && host_post && host_terminals, B200DQN_EINVAL,
             "net_train: null argument");
  for (int i = 0; i < n->nb; ++i)
    B2_REQUIRE(host_actions[i] < n->A, B200DQN_EINVAL, "net_train: action %d >= num_actions %d", host_actions[i], n->A);
  DeviceGuard g(n->device);
  cudaStream_t st = as_stream(stream);
  const size_t sb = size_t(n->nb) * kHist * kFrameBytes;
  uint8_t* p = n->h_pin;
  memcpy(p, host_pre, sb);
  memcpy(p + sb, host_post, sb);
  uint8_t* meta = p + 2 * sb;
  memcpy(meta, host_rewards, n->nb * 8);
  memcpy(meta + n->nb * 8, host_actions, n->nb);
  memcpy(meta + n->nb * 9, host_terminals, n->nb);
  B2_CHECK_CUDA(cudaMemcpyAsync(n->d_pre, p, sb, cudaMemcpyHostToDevice, st));
  B2_CHECK_CUDA(cudaMemcpyAsync(n->d_post, p + sb, sb, cudaMemcpyHostToDevice, st));
  B2_CHECK_CUDA(cudaMemcpyAsync(n->d_rew, meta, n->nb * 8, cudaMemcpyHostToDevice, st));
  B2_CHECK_CUDA(cudaMemcpyAsync(n->d_act, meta + n->nb * 8, n->nb, cudaMemcpyHostToDevice, st));
  B2_CHECK_CUDA(cudaMemcpyAsync(n->d_term, meta + n->nb * 9, n->nb, cudaMemcpyHostToDevice, st));
  int rc = b200dqn_net_train_device(n, n->d_pre, n->d_act, n->d_rew, n->d_post, n->d_term, stream);
  if (rc) return rc;
  if (host_cost) return b200dqn_net_read_costs(n, 1, host_cost, stream);
  B2_CHECK_CUDA(cudaStreamSynchronize(st));
  return B200DQN_OK;
}

static int check_fusable(b200dqn_net* n, b200dqn_replay* r) {
  B2_REQUIRE(r->device == n->device, B200DQN_EINVAL, "train_fused: replay and net live on different devices");
  B2_REQUIRE(r->h == kFrameH && r->w == kFrameW && r->hist == kHist, B200DQN_EINVAL,
             "train_fused: replay geometry differs from the network's");
  B2_REQUIRE(r->batch == n->nb * n->world, B200DQN_EINVAL,
             "train_fused: replay batch (%d) must equal the global minibatch %d x %d", r->batch, n->nb, n->world);
  return B200DQN_OK;
}

static int train_on_ring(b200dqn_net* n, b200dqn_replay* r, cudaStream_t st) {
  const int32_t* my_idx = r->d_idx + n->rank * n->nb;  // this rank's slice of the global minibatch
  // prestates = frames index-4 .. index-1, poststates = index-3 .. index (src/replay_memory.py:71-72)
  FrameSource fs{{r->d_screens, r->d_screens}, {my_idx, my_idx}, {-kHist, -kHist + 1}, {r->size, r->size}};
  n->step_replay = r;
  const int rc = train_step(n, fs, r->d_actions, r->d_rewards, r->d_terminals, my_idx, st);
  n->step_replay = nullptr;
  return rc;
}

// train_on_ring through a cached CUDA graph (the sampler is NOT part of it: the indexes are already in r)
static int train_sampled_launch(b200dqn_net* n, b200dqn_replay* r, cudaStream_t st) {
  const bool use_graph = n->use_graph && !g_prof_on && st != nullptr;
  if (!use_graph) return train_on_ring(n, r, st);
  if (!n->graph_train_exec || n->graph_train_replay != r || n->graph_train_stream != st ||
      n->graph_train_world != n->world || n->graph_train_gen != g_ktrace_gen) {
    if (n->graph_train_exec) { cudaGraphExecDestroy(n->graph_train_exec); n->graph_train_exec = nullptr; }
    cudaGraph_t graph = nullptr;
    B2_CHECK_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = train_on_ring(n, r, st);
    cudaError_t e = cudaStreamEndCapture(st, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    B2_CHECK_CUDA(e);
    B2_CHECK_CUDA(cudaGraphInstantiate(&n->graph_train_exec, graph, 0));
    cudaGraphDestroy(graph);
    n->graph_train_replay = r; n->graph_train_stream = st; n->graph_train_world = n->world;
    n->graph_train_gen = g_ktrace_gen;
  }
  B2_CHECK_CUDA(cudaGraphLaunch(n->graph_train_exec, st));
  return B200DQN_OK;
}

extern "C" int b200dqn_net_train_sampled(b200dqn_net* n, b200dqn_replay* r, void* stream) {
  B2_REQUIRE(n && r, B200DQN_EINVAL, "net_train_sampled: null argument");
  int rc = check_fusable(n, r);
  if (rc) return rc;
  DeviceGuard g(n->device);
  B2_TRY(replay_flush(r, as_stream(stream)));   // pending add()s reach the ring first (not part of the graph)
  B2_TRY(train_sampled_launch(n, r, as_stream(stream)));
  n->train_iterations += 1;
  return B200DQN_OK;
}

// wait for train step number `want` (1-based count of steps run by this net) and fetch its cost from the mapped block
static int wait_step_cost(b200dqn_net* n, cudaStream_t st, uint32_t want, float* cost) {
  int rc = poll_mapped_seq(n->h_res, want, st, "train step result");
  if (rc) return rc;
  B2_REQUIRE(n->h_res[1] == 0, B200DQN_EINVAL,
             "train: a sampled action is >= num_actions %d (IndexError at deepqnetwork.py:141 in the reference)", n->A);
  if (cost) {
    const uint32_t bits = n->h_res[4 + (want - 1) % kHostCosts];
    memcpy(cost, &bits, sizeof(float));
  }
  return B200DQN_OK;
}

extern "C" int b200dqn_net_train_sampled_cost(b200dqn_net* n, b200dqn_replay* r, float* host_cost, void* stream) {
  B2_REQUIRE(host_cost, B200DQN_EINVAL, "net_train_sampled_cost: null argument");
  int rc = b200dqn_net_train_sampled(n, r, stream);
  if (rc) return rc;
  return wait_step_cost(n, as_stream(stream), uint32_t(n->train_iterations), host_cost);
}

extern "C" int b200dqn_net_train_fused(b200dqn_net* n, b200dqn_replay* r, int nsteps, void* stream) {
  B2_REQUIRE(n && r && nsteps >= 1, B200DQN_EINVAL, "net_train_fused: bad argument");
  int rc = check_fusable(n, r);
  if (rc) return rc;
  B2_REQUIRE(r->count > r->hist, B200DQN_ESTATE, "getMinibatch: count must exceed history_length");
  B2_REQUIRE(r->rng_set, B200DQN_ESTATE, "net_train_fused: call b200dqn_replay_set_rng first");
  DeviceGuard g(n->device);
  cudaStream_t st = as_stream(stream);
  B2_TRY(replay_flush(r, st));
  // The whole step (sampler + 15 kernels, three side branches) is captured once into a CUDA graph
  // and replayed: one graph launch per step instead of ~17 stream operations.
  const bool use_graph = n->use_graph && !g_prof_on && st != nullptr;
  if (use_graph) {
    if (n->graph_replay != r || n->graph_stream != st || n->graph_world != n->world ||
        n->graph_trace_gen != g_ktrace_gen) {
      destroy_step_graphs(n);
      n->graph_replay = r; n->graph_stream = st; n->graph_world = n->world; n->graph_trace_gen = g_ktrace_gen;
    }
    // B200DQN_DEFER_FC1=1 (experiment, off by default): several steps in one call -> the fc1 update of step t rides under
    // the forward of step t+1 (net.cuh).  Parity-clean (profiles/r2p_pytest.log) but slower: the 45 MB the update moves
    // through L2 stretch whatever runs beside it, and the forward convolutions lose more (conv1 9.2 -> 13.9 us) than
    // the dgrad chain gains; 75.9 (83.2 with the driver's carveouts) vs 71.7 us per step.
    static const bool defer_on = getenv("B200DQN_DEFER_FC1") && atoi(getenv("B200DQN_DEFER_FC1")) != 0;
    const bool deferred = nsteps >= 2 && defer_on && n->use_branches && n->cfg.math_mode == B200DQN_MATH_TCGEN05 &&
                          (n->world == 1 || comm_gather_active(n, st));
    cudaGraphExec_t* exec = deferred ? &n->graph_def_exec : &n->graph_exec;
    if (!*exec) {
      cudaGraph_t graph = nullptr;
      B2_CHECK_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      const long long launches_before = g_launch_count;
      n->defer_fc1 = deferred;
      {
        const bool prev = g_pdl_suppressed;
        if (ktrace_tick(st)) g_pdl_suppressed = true;   // the sampler must not start ahead of the tick
        rc = deferred ? fc1_update_leading(n, st) : B200DQN_OK;
        if (!rc) rc = launch_sample(r, st);
        g_pdl_suppressed = prev;
      }
      if (!rc) rc = train_on_ring(n, r, st);
      n->defer_fc1 = false;
      (deferred ? n->graph_def_launches : n->graph_launches) = int(g_launch_count - launches_before);
      cudaError_t e = cudaStreamEndCapture(st, &graph);
      if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
      B2_CHECK_CUDA(e);
      B2_CHECK_CUDA(cudaGraphInstantiate(exec, graph, 0));
      cudaGraphDestroy(graph);
    }
    for (int i = 0; i < nsteps; ++i) B2_CHECK_CUDA(cudaGraphLaunch(*exec, st));
    if (deferred) {
      n->fc1_pending = true;
      B2_TRY(fc1_update_flush(n, st));
    }
  } else {
    for (int i = 0; i < nsteps; ++i) {
      const bool prev = g_pdl_suppressed;
      if (ktrace_tick(st)) g_pdl_suppressed = true;
      rc = launch_sample(r, st);
      g_pdl_suppressed = prev;
      if (rc) return rc;
      if ((rc = train_on_ring(n, r, st))) return rc;
    }
  }
  n->train_iterations += nsteps;
  r->samples_launched += uint32_t(nsteps);
  return B200DQN_OK;
}

// agent.py:102-114 as ONE call (see include/b200dqn.h)
extern "C" int b200dqn_net_step_host(b200dqn_net* n, b200dqn_replay* r, int nframes, const uint8_t* host_actions,
                                     const int64_t* host_rewards, const uint8_t* host_frames,
                                     const uint8_t* host_terminals, int train_repeat, const uint32_t* host_key624,
                                     uint32_t host_pos, float* host_costs, uint32_t* host_words_consumed, void* stream) {
  B2_REQUIRE(n && r && nframes >= 0 && train_repeat >= 0 && train_repeat <= kHostCosts, B200DQN_EINVAL,
             "net_step_host: bad argument");
  B2_REQUIRE(nframes == 0 || (host_actions && host_rewards && host_frames && host_terminals), B200DQN_EINVAL,
             "net_step_host: null frame arrays");
  for (int i = 0; i < nframes; ++i) {     // ReplayMemory.add x nframes (replay_memory.py:26-34): pinned bank, deferred
    int rc = b200dqn_replay_add(r, host_actions[i], host_rewards[i], host_frames + size_t(i) * r->frame_bytes,
                                host_terminals[i], stream);
    if (rc) return rc;
  }
  if (train_repeat == 0) return B200DQN_OK;
  DeviceGuard g(n->device);
  cudaStream_t st = as_stream(stream);
  int rc;
  if (host_key624 && (rc = replay_set_rng_async(r, host_key624, host_pos, st))) return rc;   // the host stream moved
  uint32_t words_before = 0;
  if (host_words_consumed) {              // running totals: exact across several samplings, once earlier ones are in
    if ((rc = replay_wait_words(r, st))) return rc;
    words_before = r->h_words[2];
  }
  rc = b200dqn_net_train_fused(n, r, train_repeat, stream);
  if (rc) return rc;
  if (!host_costs && !host_words_consumed) return B200DQN_OK;      // asynchronous
  float last = 0.f;
  rc = wait_step_cost(n, st, uint32_t(n->train_iterations), &last);   // the last step's cost is published last
  if (rc) return rc;
  if (host_costs)
    for (int i = 0; i < train_repeat; ++i) {
      const uint32_t bits = n->h_res[4 + (uint32_t(n->train_iterations) - train_repeat + i) % kHostCosts];
      memcpy(host_costs + i, &bits, sizeof(float));
    }
  if (host_words_consumed) {
    rc = replay_wait_words(r, st);
    if (rc) return rc;
    *host_words_consumed = r->h_words[2] - words_before;   // running totals: exact across several samplings
  }
  return B200DQN_OK;
}

extern "C" int b200dqn_net_read_costs(b200dqn_net* n, int count, float* host_costs, void* stream) {
  B2_REQUIRE(n && host_costs && count >= 1 && count <= kCostRing, B200DQN_EINVAL, "net_read_costs: bad argument");
  DeviceGuard g(n->device);
  cudaStream_t st = as_stream(stream);
  float* ring = reinterpret_cast<float*>(n->h_pin);
  uint32_t step = 0;
  // the pinned block is reused: wait for anything in flight first
  B2_CHECK_CUDA(cudaStreamSynchronize(st));
  B2_CHECK_CUDA(cudaMemcpyAsync(ring, n->d_cost, (kCostRing + 2) * sizeof(float), cudaMemcpyDeviceToHost, st));
  B2_CHECK_CUDA(cudaMemcpyAsync(&step, n->d_step, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  B2_CHECK_CUDA(cudaStreamSynchronize(st));
  B2_REQUIRE(reinterpret_cast<uint32_t*>(ring)[kCostRing + 1] == 0, B200DQN_EINVAL,
             "train: a sampled action is >= num_actions %d (IndexError at deepqnetwork.py:141 in the reference)", n->A);
  B2_REQUIRE(uint32_t(count) <= step, B200DQN_ESTATE, "net_read_costs: only %u steps have run", step);
  for (int i = 0; i < count; ++i) host_costs[i] = ring[(step - count + i) % kCostRing];
  return B200DQN_OK;
}

extern "C" int b200dqn_net_train_iterations(const b200dqn_net* n, int64_t* iters) {
  B2_REQUIRE(n && iters, B200DQN_EINVAL, "null argument");
  *iters = n->train_iterations;
  return B200DQN_OK;
}

extern "C" int b200dqn_net_device_ptr(b200dqn_net* n, int which, void** dev_ptr, size_t* bytes) {
  B2_REQUIRE(n && dev_ptr, B200DQN_EINVAL, "net_device_ptr: null argument");
  void* p = nullptr;
  size_t b = 0;
  switch (which) {
    case B200DQN_NET_PTR_Q_ONLINE: p = n->d_q[0]; b = size_t(n->nb) * n->A * 4; break;
    case B200DQN_NET_PTR_Q_TARGET: p = n->d_q[1]; b = size_t(n->nb) * n->A * 4; break;
    case B200DQN_NET_PTR_DELTAS: p = n->d_delta; b = size_t(n->nb) * n->A * 4; break;
    case B200DQN_NET_PTR_GRADS: p = n->d_g; b = n->n_params * 4; break;
    case B200DQN_NET_PTR_WEIGHTS: p = n->d_w; b = n->n_params * 4; break;
    case B200DQN_NET_PTR_COST: p = n->d_cost; b = kCostRing * 4; break;
    case B200DQN_NET_PTR_H1: p = n->d_h1[0]; b = size_t(n->nb) * kP1 * kP1 * kC1 * 4; break;
    case B200DQN_NET_PTR_H2: p = n->d_h2[0]; b = size_t(n->nb) * kP2 * kP2 * kC2 * 4; break;
    case B200DQN_NET_PTR_H3: p = n->d_h3[0]; b = size_t(n->nb) * kFlat * 4; break;
    case B200DQN_NET_PTR_H4: p = n->d_h4[0]; b = size_t(n->nb) * kHidden * 4; break;
    default: B2_REQUIRE(false, B200DQN_EINVAL, "net_device_ptr: unknown selector %d", which);
  }
  *dev_ptr = p;
  if (bytes) *bytes = b;
  return B200DQN_OK;
}

extern "C" int b200dqn_net_set_keep_grads(b200dqn_net* n, int keep) {
  B2_REQUIRE(n, B200DQN_EINVAL, "null net");
  n->keep_grads = keep != 0;
  destroy_step_graphs(n);   // parameters are baked in
  if (n->graph_train_exec) { cudaGraphExecDestroy(n->graph_train_exec); n->graph_train_exec = nullptr; }
  return B200DQN_OK;
}

extern "C" int b200dqn_net_get_grads(b200dqn_net* n, int layer, float* host_dW, void* stream) {
  B2_REQUIRE(n && host_dW && layer >= 0 && layer < kLayers, B200DQN_EINVAL, "net_get_grads: bad argument");
  B2_REQUIRE(n->cfg.math_mode != B200DQN_MATH_TCGEN05 || n->world > 1 || n->keep_grads ||
                 getenv("B200DQN_FC1_FUSED") == nullptr,
             B200DQN_ESTATE, "net_get_grads: call b200dqn_net_set_keep_grads(net, 1) before the step (fused optimizer)");
  DeviceGuard g(n->device);
  cudaStream_t st = as_stream(stream);
  const int64_t n4 = n->n_params / 4;
  if (n->world == 1) {  // partials of the last step are still in scratch; sum them into d_g
    k_optimizer<<<cdiv(n4, 256), 256, 0, st>>>(n->lt, n->d_part, n->d_g, n->d_w, n->d_s, 0, n4, 1 | 2, OptArgs{},
                                               KTrace{nullptr, 0});
    B2_LAUNCH_CHECK();
  } else if (n->xchg_ok && n->xchg_sched == 2 && n->d_xbuf && layer == 3) {
    // gather schedule: fc1's global gradient was computed locally and never passed through d_g
    const int64_t b4 = n->lt.off[3] / 4, e4 = n->lt.off[4] / 4;
    k_optimizer<<<cdiv(e4 - b4, 256), 256, 0, st>>>(n->lt, n->d_part, n->d_g, n->d_w, n->d_s, b4, e4, 1 | 2, OptArgs{},
                                                    KTrace{nullptr, 0});
    B2_LAUNCH_CHECK();
  }
  return xfer_params(n, n->d_g, layer, host_dW, false, st);
}

extern "C" int b200dqn_net_launches_per_step(const b200dqn_net* n, int* launches) {
  B2_REQUIRE(n && launches, B200DQN_EINVAL, "null argument");
  // Counted at the launch sites while the step was captured into its CUDA graph; before the first
  // fused step: the static schedule (sample, 4 forward, head, 7 backward GEMMs, per-layer optimizers).
  if (n->graph_def_launches > 0) {          // the multi-step graph (one gated fc1 update at its head)
    *launches = n->graph_def_launches;
  } else if (n->graph_launches > 0) {
    *launches = n->graph_launches;
  } else {
    const bool tc = n->cfg.math_mode == B200DQN_MATH_TCGEN05;
    *launches = 1 + 4 + 1 + 7 + (n->world > 1 ? (tc ? 7 : 2) : (tc ? 6 : 4));
  }
  return B200DQN_OK;
}
