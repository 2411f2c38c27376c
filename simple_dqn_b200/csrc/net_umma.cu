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

// ------------------------------------------------------------------------------------------
// Backward problems (online network).  dZ tensors carry the Rectlin mask of their producer.
// ------------------------------------------------------------------------------------------

// fc1 dgrad, swapped: dZ3[b][m] = (sum_n W4[m][n] * dZ4[b][n]) * (H3[b][m] > 0);  M = 3136, N = rows, K = 512
struct UFc1Dgrad {
  static constexpr int kBN = 32;
  static constexpr bool kAExact = false, kARowMajorThreads = true, kBRowMajorThreads = true;
  const float* w4;   // [3136][512]
  const float* dz4;  // [rows][512]
  const float* h3;   // [rows][3136]
  float* dz3;        // [rows][3136]
  int rows;
  __device__ int M(int) const { return kFlat; }
  __device__ int N(int) const { return rows; }
  __device__ void krange(int, int& kb, int& ke) const { kb = 0; ke = kHidden / 64; }
  __device__ void a8(int, int m, int k0, float v[8]) const {
    if (m >= kFlat) { zero8(v); return; }
    ld8(w4 + m * kHidden + k0, v);
  }
  __device__ void b8(int, int n, int k0, float v[8]) const {
    if (n >= rows) { zero8(v); return; }
    ld8(dz4 + n * kHidden + k0, v);
  }
  __device__ void store8(int, int m, int n0, const float v[8]) const {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (n0 + j < rows) {
        const int i = (n0 + j) * kFlat + m;
        dz3[i] = h3[i] > 0.f ? v[j] : 0.f;
      }
  }
};

// fc1 wgrad: dW4[m][n] = sum_b H3[b][m] * dZ4[b][n];  M = 3136, N = 512 (tiles of 128), K = rows (padded to 64)
struct UFc1Wgrad {
  static constexpr int kBN = 128;
  static constexpr bool kAExact = false, kARowMajorThreads = false, kBRowMajorThreads = false;
  const float* h3;
  const float* dz4;
  float* dw4;  // [3136][512]
  int rows;
  __device__ int M(int) const { return kFlat; }
  __device__ int N(int) const { return kHidden; }
  __device__ void krange(int, int& kb, int& ke) const { kb = 0; ke = (rows + 63) / 64; }
  __device__ void a8(int, int m, int k0, float v[8]) const {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (m < kFlat && k0 + j < rows) ? h3[(k0 + j) * kFlat + m] : 0.f;
  }
  __device__ void b8(int, int n, int k0, float v[8]) const {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (k0 + j < rows) ? dz4[(k0 + j) * kHidden + n] : 0.f;
  }
  __device__ void store8(int, int m, int n0, const float v[8]) const { st8(dw4 + m * kHidden + n0, v); }
};

// conv dgrad by output-parity class z (see ConvDgrad in net_simt.cuh):
//   dX[n,y,x,c] = (sum_{r',s',ko} dZ[n, yy-r', xx-s', ko] * W[(r,s,c)][ko]) * (X > 0)
template <int H, int C, int R, int ST, int KO>
struct UConvDgrad {
  static constexpr int P = (H - R) / ST + 1, RT = R / ST, HC = (H + ST - 1) / ST, K = RT * RT * KO;
  static_assert(K % 64 == 0 && KO % 8 == 0, "k-blocks of 64");
  static constexpr int kBN = C;
  static constexpr bool kAExact = false, kARowMajorThreads = true, kBRowMajorThreads = true;
  const float* dz;  // [rows][P][P][KO]
  const float* w;   // [(r,s,c)][KO]
  const float* x;   // [rows][H][H][C]
  float* dx;
  int rows;
  __device__ int M(int) const { return rows * HC * HC; }
  __device__ int N(int) const { return C; }
  __device__ void krange(int, int& kb, int& ke) const { kb = 0; ke = K / 64; }
  __device__ void a8(int z, int m, int k0, float v[8]) const {
    zero8(v);
    if (m >= rows * HC * HC) return;
    const int n = m / (HC * HC), yx = m % (HC * HC), yy = yx / HC, xx = yx % HC;
    const int rp = k0 / (RT * KO), sp = (k0 / KO) % RT, ko = k0 % KO;
    if (yy * ST + z / ST >= H || xx * ST + z % ST >= H) return;
    const int p = yy - rp, q = xx - sp;
    if (p < 0 || p >= P || q < 0 || q >= P) return;
    ld8(dz + ((n * P + p) * P + q) * KO + ko, v);
  }
  __device__ void b8(int z, int c, int k0, float v[8]) const {
    const int rp = k0 / (RT * KO), sp = (k0 / KO) % RT, ko = k0 % KO;
    const int r = rp * ST + z / ST, s = sp * ST + z % ST;
    ld8(w + ((r * R + s) * C + c) * KO + ko, v);
  }
  __device__ void store8(int z, int m, int c0, const float v[8]) const {
    const int n = m / (HC * HC), yx = m % (HC * HC), yy = yx / HC, xx = yx % HC;
    const int y = yy * ST + z / ST, xq = xx * ST + z % ST;
    if (y >= H || xq >= H) return;
    const int i = ((n * H + y) * H + xq) * C + c0;
    float xv[8], o[8];
    ld8(x + i, xv);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = xv[j] > 0.f ? v[j] : 0.f;
    st8(dx + i, o);
  }
};

// conv wgrad with split-K over z: part[z][(r,s,c)][ko] = sum_{(n,p,q) in split} X[n,p*ST+r,q*ST+s,c] * dZ[n,p,q,ko]
template <int H, int C, int R, int ST, int KO>
struct UConvWgrad {
  static constexpr int P = (H - R) / ST + 1, KW = R * R * C;
  static constexpr int kBN = KO;
  static constexpr bool kAExact = false, kARowMajorThreads = false, kBRowMajorThreads = false;
  const float* x;   // [rows][H][H][C]
  const float* dz;  // [rows][P][P][KO]
  float* part;      // [splits][KW][KO]
  int rows, kb_per_split;
  __device__ int M(int) const { return KW; }
  __device__ int N(int) const { return KO; }
  __device__ void krange(int z, int& kb, int& ke) const {
    const int total = (rows * P * P + 63) / 64;
    kb = z * kb_per_split;
    ke = min(kb + kb_per_split, total);
  }
  __device__ void a8(int, int m, int k0, float v[8]) const {
    const int r = m / (R * C), sc = m % (R * C);
    int n = k0 / (P * P), pq = k0 % (P * P), p = pq / P, q = pq % P;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = (m < KW && n < rows) ? x[((n * H + p * ST + r) * H + q * ST) * C + sc] : 0.f;
      if (++q == P) { q = 0; if (++p == P) { p = 0; ++n; } }
    }
  }
  __device__ void b8(int, int ko, int k0, float v[8]) const {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (k0 + j < rows * P * P) ? dz[(k0 + j) * KO + ko] : 0.f;
  }
  __device__ void store8(int z, int m, int n0, const float v[8]) const { st8(part + (z * KW + m) * KO + n0, v); }
};

// conv1 wgrad: A = u8 frame pixels (exact in fp16), m = (c, r, s); the 1/255 goes on the store.
struct UConv1Wgrad {
  static constexpr int kBN = 32;
  static constexpr bool kAExact = true, kARowMajorThreads = false, kBRowMajorThreads = false;
  const uint8_t* src;
  const int32_t* idx;
  int shift;
  const float* dz;  // dZ1 [rows][20][20][32]
  float* part;      // [splits][256][32]
  int rows, kb_per_split;
  __device__ int M(int) const { return kK1; }
  __device__ int N(int) const { return kC1; }
  __device__ void krange(int z, int& kb, int& ke) const {
    const int total = (rows * kP1 * kP1 + 63) / 64;
    kb = z * kb_per_split;
    ke = min(kb + kb_per_split, total);
  }
  __device__ void a8(int, int m, int k0, float v[8]) const {
    const int c = m >> 6, r = (m >> 3) & 7, s = m & 7;
    int n = k0 / (kP1 * kP1), pq = k0 % (kP1 * kP1), p = pq / kP1, q = pq % kP1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float val = 0.f;
      if (n < rows) {
        const int64_t f = static_cast<int64_t>(idx[n]) + shift + c;
        val = float(src[f * kFrameBytes + (p * 4 + r) * kFrameW + q * 4 + s]);
      }
      v[j] = val;
      if (++q == kP1) { q = 0; if (++p == kP1) { p = 0; ++n; } }
    }
  }
  __device__ void b8(int, int ko, int k0, float v[8]) const {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (k0 + j < rows * kP1 * kP1) ? dz[(k0 + j) * kC1 + ko] : 0.f;
  }
  __device__ void store8(int z, int m, int n0, const float v[8]) const {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = v[j] * (1.0f / 255.0f);
    st8(part + (z * kK1 + m) * kC1 + n0, o);
  }
};

constexpr int kUFc1Splits = 7;    // 49 k-blocks of 64 -> 7 per CTA; 4 M-tiles x 7 x 2 nets = 56 CTAs
constexpr int kUWgradKb = 4;      // k-blocks (256 pixels) per wgrad split

int umma_wgrad_splits(int layer, int rows) {
  const int kred = layer == 0 ? rows * kP1 * kP1 : layer == 1 ? rows * kP2 * kP2 : rows * kP3 * kP3;
  const int kbs = (kred + 63) / 64;
  return (kbs + kUWgradKb - 1) / kUWgradKb;
}

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

int umma_backward_op(b200dqn_net* n, int op, const uint8_t* src, const int32_t* idx, int shift, int rows,
                     cudaStream_t st) {
  const LayerTable& lt = n->lt;
  const float* w = n->d_w;
  switch (op) {
    case 0: {
      UFc1Wgrad p{n->d_h3[0], n->d_dz4, n->d_part + lt.part_off[3], rows};
      return umma::launch_umma("fc1_wgrad", p, kFlat, kHidden, 1, st);
    }
    case 1: {
      UFc1Dgrad p{w + lt.off[3], n->d_dz4, n->d_h3[0], n->d_dz3, rows};
      return umma::launch_umma("fc1_dgrad", p, kFlat, rows, 1, st);
    }
    case 2: {
      using P = UConvWgrad<kP2, kC2, 3, 1, kC3>;
      P p{n->d_h2[0], n->d_dz3, n->d_part + lt.part_off[2], rows, kUWgradKb};
      return umma::launch_umma("conv3_wgrad", p, P::KW, kC3, lt.splits[2], st);
    }
    case 3: {
      using P = UConvDgrad<kP2, kC2, 3, 1, kC3>;
      P p{n->d_dz3, w + lt.off[2], n->d_h2[0], n->d_dz2, rows};
      return umma::launch_umma("conv3_dgrad", p, rows * P::HC * P::HC, kC2, 1, st);
    }
    case 4: {
      using P = UConvWgrad<kP1, kC1, 4, 2, kC2>;
      P p{n->d_h1[0], n->d_dz2, n->d_part + lt.part_off[1], rows, kUWgradKb};
      return umma::launch_umma("conv2_wgrad", p, P::KW, kC2, lt.splits[1], st);
    }
    case 5: {
      using P = UConvDgrad<kP1, kC1, 4, 2, kC2>;
      P p{n->d_dz2, w + lt.off[1], n->d_h1[0], n->d_dz1, rows};
      return umma::launch_umma("conv2_dgrad", p, rows * P::HC * P::HC, kC1, 4, st);
    }
    default: {
      UConv1Wgrad p{src, idx, shift, n->d_dz1, n->d_part + lt.part_off[0], rows, kUWgradKb};
      return umma::launch_umma("conv1_wgrad", p, kK1, kC1, lt.splits[0], st);
    }
  }
}

int umma_fc1_splits() { return kUFc1Splits; }
bool umma_has_backward() { return true; }
int umma_forward_launches() { return 4; }
int umma_backward_launches() { return 7; }

}  // namespace b200
