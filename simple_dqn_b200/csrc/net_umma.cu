// net_umma.cu — math_mode TCGEN05: the GEMM-shaped ops of the Nature-DQN step on the 5th-gen
// tensor cores (tcgen05.mma, TMEM accumulators) through the software-staged implicit-GEMM kernel
// of umma.cuh.  Same HBM buffers and layouts as the fp32 SIMT engine (net_simt.cuh), so every
// kernel here is checked against its SIMT twin and against the CPU oracle.
#include "net.cuh"
#include "net_umma.cuh"
#include "umma.cuh"

namespace b200 {

__device__ __forceinline__ void ld8(const float* p, float v[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float v[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void zero8(float v[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
}

// ------------------------------------------------------------------------------------------
// Forward problems (z = network: 0 online / prestates, 1 target / poststates)
// ------------------------------------------------------------------------------------------

// conv1: M = rows*400 output pixels, N = 32, K = 256 = (c, r, s).  One 8-element K chunk is one
// 8-pixel filter row of one frame: 8 contiguous bytes of the ring.  u8 is exact in fp16 -> no A_lo.
struct UConv1Fwd {
  static constexpr int kBN = 32;
  static constexpr bool kAExact = true, kARowMajorThreads = true, kBRowMajorThreads = false;
  const uint8_t* src[2];
  const int32_t* idx[2];
  int shift[2];
  const float* w[2];   // [256][32]
  float* out[2];       // [rows*400][32]
  int rows;
  __device__ int M(int) const { return rows * kP1 * kP1; }
  __device__ int N(int) const { return kC1; }
  __device__ void krange(int, int& kb, int& ke) const { kb = 0; ke = kK1 / 64; }
  __device__ void a8(int z, int m, int k0, float v[8]) const {
    if (m >= rows * kP1 * kP1) { zero8(v); return; }
    const int n = m / (kP1 * kP1), pq = m % (kP1 * kP1), p = pq / kP1, q = pq % kP1;
    const int c = k0 >> 6, r = (k0 >> 3) & 7;
    const int64_t f = static_cast<int64_t>((z ? idx[1] : idx[0])[n]) + (z ? shift[1] : shift[0]) + c;
    const uint8_t* ptr = (z ? src[1] : src[0]) + f * kFrameBytes + (p * 4 + r) * kFrameW + q * 4;
    const uint32_t lo = *reinterpret_cast<const uint32_t*>(ptr), hi = *reinterpret_cast<const uint32_t*>(ptr + 4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = float((lo >> (8 * j)) & 0xffu);
      v[4 + j] = float((hi >> (8 * j)) & 0xffu);
    }
  }
  __device__ void b8(int z, int n, int k0, float v[8]) const {
    const float* ww = z ? w[1] : w[0];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = ww[(k0 + j) * kC1 + n];
  }
  __device__ void store8(int z, int m, int n0, const float v[8]) const {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaxf(v[j] * (1.0f / 255.0f), 0.f);   // /255 of _setInput + Rectlin
    st8((z ? out[1] : out[0]) + m * kC1 + n0, o);
  }
};

// conv2 / conv3: NHWC fp32 input, K = (r, s, c): an 8-element chunk is 8 contiguous channels.
template <int H, int C, int R, int ST, int KO>
struct UConvFwd {
  static constexpr int P = (H - R) / ST + 1, K = R * R * C;
  static_assert(K % 64 == 0 && C % 8 == 0, "k-blocks of 64, chunks of 8 channels");
  static constexpr int kBN = KO;
  static constexpr bool kAExact = false, kARowMajorThreads = true, kBRowMajorThreads = false;
  const float* in[2];
  const float* w[2];
  float* out[2];
  int rows;
  __device__ int M(int) const { return rows * P * P; }
  __device__ int N(int) const { return KO; }
  __device__ void krange(int, int& kb, int& ke) const { kb = 0; ke = K / 64; }
  __device__ void a8(int z, int m, int k0, float v[8]) const {
    if (m >= rows * P * P) { zero8(v); return; }
    const int n = m / (P * P), pq = m % (P * P), p = pq / P, q = pq % P;
    const int r = k0 / (R * C), sc = k0 % (R * C);
    ld8((z ? in[1] : in[0]) + ((n * H + p * ST + r) * H + q * ST) * C + sc, v);
  }
  __device__ void b8(int z, int n, int k0, float v[8]) const {
    const float* ww = z ? w[1] : w[0];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = ww[(k0 + j) * KO + n];
  }
  __device__ void store8(int z, int m, int n0, const float v[8]) const {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaxf(v[j], 0.f);
    st8((z ? out[1] : out[0]) + m * KO + n0, o);
  }
};

// fc1 forward, operands swapped so the 512 hidden units fill the UMMA M dimension:
//   part[z][b][m] = sum_{k in split} W4[k][m] * H3[b][k],   z = net * splits + split
struct UFc1Fwd {
  static constexpr int kBN = 32;
  static constexpr bool kAExact = false, kARowMajorThreads = false, kBRowMajorThreads = true;
  const float* in[2];   // H3 [rows][3136]
  const float* w[2];    // W4 [3136][512]
  float* part;          // [2*splits][rows][512]
  int rows, splits;
  __device__ int M(int) const { return kHidden; }
  __device__ int N(int) const { return rows; }
  __device__ void krange(int z, int& kb, int& ke) const {
    const int per = (kFlat / 64 + splits - 1) / splits;
    kb = (z % splits) * per;
    ke = min(kb + per, kFlat / 64);
  }
  __device__ void a8(int z, int m, int k0, float v[8]) const {
    const float* ww = (z / splits) ? w[1] : w[0];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = ww[(k0 + j) * kHidden + m];
  }
  __device__ void b8(int z, int n, int k0, float v[8]) const {
    if (n >= rows) { zero8(v); return; }
    ld8(((z / splits) ? in[1] : in[0]) + n * kFlat + k0, v);
  }
  __device__ void store8(int z, int m, int n0, const float v[8]) const {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (n0 + j < rows) part[(z * rows + n0 + j) * kHidden + m] = v[j];
  }
};

constexpr int kUFc1Splits = 7;   // 49 k-blocks of 64 -> 7 per CTA; 4 M-tiles x 7 x 2 nets = 56 CTAs

int umma_net_init(b200dqn_net* n) {
  (void)n;
  return B200DQN_OK;
}
void umma_net_destroy(b200dqn_net*) {}
int umma_weights_changed(b200dqn_net*, cudaStream_t) { return B200DQN_OK; }
int umma_target_synced(b200dqn_net*, cudaStream_t) { return B200DQN_OK; }

int umma_forward(b200dqn_net* n, const uint8_t* const src[2], const int32_t* const idx[2], const int shift[2],
                 int nets, int rows, cudaStream_t st) {
  const LayerTable& lt = n->lt;
  const float* w[2] = {n->d_w, n->d_tw};
  int rc;
  {
    UConv1Fwd p;
    for (int z = 0; z < 2; ++z) {
      p.src[z] = src[z]; p.idx[z] = idx[z]; p.shift[z] = shift[z];
      p.w[z] = w[z] + lt.off[0]; p.out[z] = n->d_h1[z];
    }
    p.rows = rows;
    if ((rc = umma::launch_umma("conv1_fwd", p, rows * kP1 * kP1, kC1, nets, st))) return rc;
  }
  {
    using P = UConvFwd<kP1, kC1, 4, 2, kC2>;
    P p;
    for (int z = 0; z < 2; ++z) { p.in[z] = n->d_h1[z]; p.w[z] = w[z] + lt.off[1]; p.out[z] = n->d_h2[z]; }
    p.rows = rows;
    if ((rc = umma::launch_umma("conv2_fwd", p, rows * kP2 * kP2, kC2, nets, st))) return rc;
  }
  {
    using P = UConvFwd<kP2, kC2, 3, 1, kC3>;
    P p;
    for (int z = 0; z < 2; ++z) { p.in[z] = n->d_h2[z]; p.w[z] = w[z] + lt.off[2]; p.out[z] = n->d_h3[z]; }
    p.rows = rows;
    if ((rc = umma::launch_umma("conv3_fwd", p, rows * kP3 * kP3, kC3, nets, st))) return rc;
  }
  {
    UFc1Fwd p;
    for (int z = 0; z < 2; ++z) { p.in[z] = n->d_h3[z]; p.w[z] = w[z] + lt.off[3]; }
    p.part = n->d_fc1part; p.rows = rows; p.splits = kUFc1Splits;
    if ((rc = umma::launch_umma("fc1_fwd", p, kHidden, rows, nets * kUFc1Splits, st))) return rc;
  }
  return B200DQN_OK;
}

int umma_fc1_splits() { return kUFc1Splits; }
bool umma_has_backward() { return false; }
int umma_backward(b200dqn_net*, const uint8_t*, const int32_t*, int, int, cudaStream_t) { return B200DQN_ENOTIMPL; }
int umma_forward_launches() { return 4; }
int umma_backward_launches() { return 7; }

}  // namespace b200
