// net_umma.cu — math_mode TCGEN05: the GEMM-shaped ops of the Nature-DQN step on the 5th-gen
// tensor cores (tcgen05.mma, TMEM accumulators): forward + dgrad through the K-major kernel of
// umma2.cuh, wgrad through the MN-major kernel of umma_mn.cuh, plus the fused optimizer / tile-image
// kernels.  Same fp32 HBM tensors as the SIMT engine (net_simt.cuh), so every kernel here is checked
// against its SIMT twin and against the CPU oracle.
#include "net.cuh"
#include "net_umma.cuh"
#include "umma.cuh"
#include "umma2.cuh"
#include "umma_mn.cuh"
#include "conv1_tma.cuh"

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

// ==========================================================================================
// Engine v2 (umma2.cuh): pre-split fp16 hi/lo operand planes + weight tile images.
// ==========================================================================================
struct UmmaState {
  // activation planes per net: [hi plane | lo plane], NHWC fp16
  __half* h16[3][2] = {};     // H1, H2, H3  x  (online, target)
  int64_t h_elems[3] = {};
  __half* dz16[4] = {};       // dZ4, dZ3, dZ2, dZ1 (online)
  int64_t dz_elems[4] = {};
  // weight tile images ([hi | lo] per (tile, k-block))
  uint8_t* img_fwd[2][4] = {};  // conv1, conv2, conv3, fc1  x  (online, target)
  int64_t img_fwd_bytes[4] = {};
  uint8_t* im2col1 = nullptr;   // conv1_fwd's A_hi tiles of the online net: [ceil(nb*400/128)][4][16 KB]
  uint8_t* img_dgr[3] = {};     // fc1_dgrad (A operand), conv3_dgrad (B), conv2_dgrad (B, 4 parity classes)
};
static inline UmmaState* ust(b200dqn_net* n) { return static_cast<UmmaState*>(n->umma_state); }

struct PlanePair {
  __half* hi;
  int64_t lo_off;   // lo plane = hi + lo_off
};

__device__ __forceinline__ void store_f32_and_planes(float* f32, const PlanePair& pl, int64_t i, const float v[8]) {
  if (f32) st8(f32 + i, v);
  umma2::split8_planes(v, pl.hi + i, pl.hi + pl.lo_off + i);
}

// ---- forward -----------------------------------------------------------------------------
struct V2Conv1Fwd {
  static constexpr int kBN = 32;
  static constexpr bool kAExact = true, kARowMajorThreads = true, kBRowMajorThreads = false;
  static constexpr int kAMode = umma2::kReg, kBMode = umma2::kBulk;
  static constexpr bool kStagedEpilogue = true, kDumpA = true, kPrefetch = false;
  uint8_t* im2col;          // online net only: [mtile][4 kb][128 x 128 B] A_hi tiles for conv1_wgrad (nullptr = off)
  const uint8_t* src[2];
  const int32_t* idx[2];
  int shift[2];
  const uint8_t* wimg[2];   // [4 kb][hi 32x128 | lo 32x128]
  float* out[2];
  PlanePair out16[2];
  int rows;
  __device__ int M(int) const { return rows * kP1 * kP1; }
  __device__ int N(int) const { return kC1; }
  __device__ void krange(int, int& kb, int& ke) const { kb = 0; ke = kK1 / 64; }
  // first byte of output pixel m's receptive field in frame 0 of its sample (nullptr = padding row)
  __device__ const uint8_t* a_row_ptr(int z, int m) const {
    if (m >= rows * kP1 * kP1) return nullptr;
    const int n = m / (kP1 * kP1), pq = m % (kP1 * kP1), p = pq / kP1, q = pq % kP1;
    const int64_t f = static_cast<int64_t>((z ? idx[1] : idx[0])[n]) + (z ? shift[1] : shift[0]);
    return (z ? src[1] : src[0]) + f * kFrameBytes + (p * 4) * kFrameW + q * 4;
  }
  __device__ uint2 a_raw8(const uint8_t* row, int k0) const {   // k0 = (c, r, 0): 8 pixels of filter row r, frame c
    if (!row) return make_uint2(0u, 0u);
    const int c = k0 >> 6, r = (k0 >> 3) & 7;
    const uint8_t* ptr = row + c * kFrameBytes + r * kFrameW;
    return make_uint2(*reinterpret_cast<const uint32_t*>(ptr), *reinterpret_cast<const uint32_t*>(ptr + 4));
  }
  __device__ static void cvt8(uint2 raw, float v[8]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = float((raw.x >> (8 * j)) & 0xffu);
      v[4 + j] = float((raw.y >> (8 * j)) & 0xffu);
    }
  }
  __device__ const uint8_t* b_tile(int z, int, int kb) const { return (z ? wimg[1] : wimg[0]) + kb * (kC1 * 256); }
  __device__ uint8_t* a_dump(int z, int mtile, int kb) const {
    return (z == 0 && im2col) ? im2col + (int64_t(mtile) * (kK1 / 64) + kb) * (128 * 128) : nullptr;
  }
  __device__ void store8(int z, int m, int n0, const float v[8]) const {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaxf(v[j] * (1.0f / 255.0f), 0.f);
    store_f32_and_planes(z ? out[1] : out[0], z ? out16[1] : out16[0], int64_t(m) * kC1 + n0, o);
  }
};

template <int H, int C, int R, int ST, int KO>
struct V2ConvFwd {
  static constexpr int P = (H - R) / ST + 1, K = R * R * C;
  static_assert(K % 64 == 0 && C % 8 == 0, "k-blocks of 64, chunks of 8 channels");
  static constexpr int kBN = KO;
  static constexpr bool kAExact = false, kARowMajorThreads = true, kBRowMajorThreads = false;
  static constexpr int kAMode = umma2::kAsync, kBMode = umma2::kBulk;
  static constexpr bool kStagedEpilogue = true, kDumpA = false, kPrefetch = false;
  PlanePair in16[2];
  const uint8_t* wimg[2];   // [K/64][hi KOx128 | lo KOx128]
  float* out[2];
  PlanePair out16[2];
  int rows;
  __device__ int M(int) const { return rows * P * P; }
  __device__ int N(int) const { return KO; }
  __device__ void krange(int, int& kb, int& ke) const { kb = 0; ke = K / 64; }
  __device__ umma2::Planes a_planes(int z) const { return {z ? in16[1].hi : in16[0].hi, in16[0].lo_off}; }
  __device__ umma2::RowCtx a_row(int, int m) const {
    const int n = m / (P * P), pq = m % (P * P), p = pq / P, q = pq % P;
    return {(int64_t(n * H + p * ST) * H + q * ST) * C, 0, 0, m < rows * P * P};
  }
  __device__ bool a_chunk(int, const umma2::RowCtx& rc, int kk, int64_t& off) const {
    const int r = kk / (R * C), sc = kk % (R * C);
    off = rc.base + r * (H * C) + sc;
    return true;
  }
  __device__ const uint8_t* b_tile(int z, int, int kb) const { return (z ? wimg[1] : wimg[0]) + kb * (KO * 256); }
  __device__ void store8(int z, int m, int n0, const float v[8]) const {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaxf(v[j], 0.f);
    store_f32_and_planes(z ? out[1] : out[0], z ? out16[1] : out16[0], int64_t(m) * KO + n0, o);
  }
};

struct V2Fc1Fwd {
  static constexpr int kBN = 32;
  static constexpr bool kAExact = false, kARowMajorThreads = false, kBRowMajorThreads = true;
  static constexpr int kAMode = umma2::kBulk, kBMode = umma2::kAsync;
  static constexpr bool kStagedEpilogue = false, kDumpA = false, kPrefetch = false;   // strided outputs; lanes run along m
  // The weights come from the ROW-oriented fc1 image (the one fc1_dgrad reads K-major: rows = flat index, 64 hidden
  // units per 128-byte row): read with "row = k" it is an M-contiguous (MN-major) A operand, so the forward needs no
  // image of its own and nothing has to be re-packed after the optimizer.
  static constexpr bool kAMnMajor = true;
  static constexpr uint32_t kAMnLoOffset = 128 * 128;   // lo half of a [hi 128x128 B | lo 128x128 B] tile
  PlanePair in16[2];        // H3 planes [rows][3136]
  const uint8_t* wimg[2];   // [25 flat tiles][8 hidden blocks][hi 128x128 | lo 128x128]
  float* part;              // [2*splits][rows][512]
  int rows, splits;
  __device__ int M(int) const { return kHidden; }
  __device__ int N(int) const { return rows; }
  __device__ void krange(int z, int& kb, int& ke) const {
    const int per = (kFlat / 64 + splits - 1) / splits;
    kb = (z % splits) * per;
    ke = min(kb + per, kFlat / 64);
  }
  // hi sub-tile [64 flat rows x 64 hidden] of hidden block 2*mtile + chunk, flat k-block kb
  __device__ const uint8_t* a_sub(int z, int mtile, int kb, int chunk) const {
    return ((z / splits) ? wimg[1] : wimg[0]) + (int64_t(kb >> 1) * (kHidden / 64) + 2 * mtile + chunk) * (128 * 256) +
           (kb & 1) * (64 * 128);
  }
  __device__ umma2::Planes b_planes(int z) const { return {(z / splits) ? in16[1].hi : in16[0].hi, in16[0].lo_off}; }
  __device__ umma2::RowCtx b_row(int, int n) const { return {int64_t(n) * kFlat, 0, 0, n < rows}; }
  __device__ bool b_chunk(int, const umma2::RowCtx& rc, int kk, int64_t& off) const {
    off = rc.base + kk;
    return true;
  }
  __device__ void store8(int z, int m, int n0, const float v[8]) const {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (n0 + j < rows) part[(z * rows + n0 + j) * kHidden + m] = v[j];
  }
};

// ---- dgrad -------------------------------------------------------------------------------
struct V2Fc1Dgrad {
  static constexpr int kBN = 32;
  static constexpr bool kAExact = false, kARowMajorThreads = true, kBRowMajorThreads = true;
  static constexpr int kAMode = umma2::kBulk, kBMode = umma2::kAsync;
  static constexpr bool kStagedEpilogue = false, kDumpA = false, kPrefetch = true;
  const uint8_t* wimg;   // [25 mtiles][8 kb][hi | lo]   rows m = flat index (p,q,c), K = hidden unit
  PlanePair dz4;         // [rows][512]
  const float* h3;       // [rows][3136] (mask)
  float* dz3;
  PlanePair dz3_16;
  int rows;
  __device__ int M(int) const { return kFlat; }
  __device__ int N(int) const { return rows; }
  __device__ void krange(int, int& kb, int& ke) const { kb = 0; ke = kHidden / 64; }
  __device__ const uint8_t* a_tile(int, int mtile, int kb) const {
    return wimg + (int64_t(mtile) * (kHidden / 64) + kb) * (128 * 256);
  }
  __device__ umma2::Planes b_planes(int) const { return {dz4.hi, dz4.lo_off}; }
  __device__ umma2::RowCtx b_row(int, int n) const { return {int64_t(n) * kHidden, 0, 0, n < rows}; }
  __device__ bool b_chunk(int, const umma2::RowCtx& rc, int kk, int64_t& off) const {
    off = rc.base + kk;
    return true;
  }
  __device__ void prefetch8(int, int m, int n0, float pf[8]) const {   // Rectlin mask of H3
#pragma unroll
    for (int j = 0; j < 8; ++j) pf[j] = (n0 + j < rows) ? h3[int64_t(n0 + j) * kFlat + m] : 0.f;
  }
  __device__ void store8p(int, int m, int n0, const float v[8], const float pf[8]) const {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (n0 + j < rows) {
        const int64_t i = int64_t(n0 + j) * kFlat + m;
        const float o = pf[j] > 0.f ? v[j] : 0.f;
        if (dz3) dz3[i] = o;
        __half h, l;
        umma2::split1(o, h, l);
        dz3_16.hi[i] = h;
        dz3_16.hi[dz3_16.lo_off + i] = l;
      }
  }
};

template <int H, int C, int R, int ST, int KO, int STG = 0>
struct V2ConvDgrad {
  static constexpr int kStagesOverride = STG;   // 0 = the deepest ring that fits (umma2::Cfg2)
  static constexpr int P = (H - R) / ST + 1, RT = R / ST, HC = (H + ST - 1) / ST, K = RT * RT * KO;
  static_assert(K % 64 == 0 && KO % 8 == 0, "k-blocks of 64");
  static constexpr int kBN = C;
  static constexpr bool kAExact = false, kARowMajorThreads = true, kBRowMajorThreads = true;
  static constexpr int kAMode = umma2::kAsync, kBMode = umma2::kBulk;
  static constexpr bool kStagedEpilogue = true, kDumpA = false, kPrefetch = true;
  PlanePair dz;          // [rows][P][P][KO]
  const uint8_t* wimg;   // [ST*ST classes][K/64][hi Cx128 | lo Cx128]
  const float* x;        // forward activation (mask)
  float* dx;
  PlanePair dx16;        // hi == nullptr -> not needed
  int rows;
  __device__ int M(int) const { return rows * HC * HC; }
  __device__ int N(int) const { return C; }
  __device__ void krange(int, int& kb, int& ke) const { kb = 0; ke = K / 64; }
  __device__ umma2::Planes a_planes(int) const { return {dz.hi, dz.lo_off}; }
  __device__ umma2::RowCtx a_row(int z, int m) const {
    const int n = m / (HC * HC), yx = m % (HC * HC), yy = yx / HC, xx = yx % HC;
    const bool ok = m < rows * HC * HC && yy * ST + z / ST < H && xx * ST + z % ST < H;
    return {(int64_t(n * P + yy) * P + xx) * KO, yy, xx, ok};
  }
  __device__ bool a_chunk(int, const umma2::RowCtx& rc, int kk, int64_t& off) const {
    const int rp = kk / (RT * KO), sp = (kk / KO) % RT, ko = kk % KO;
    const int p = rc.y - rp, q = rc.x - sp;
    off = rc.base - (rp * P + sp) * KO + ko;
    return p >= 0 && p < P && q >= 0 && q < P;
  }
  __device__ const uint8_t* b_tile(int z, int, int kb) const { return wimg + (int64_t(z) * (K / 64) + kb) * (C * 256); }
  __device__ void prefetch8(int z, int m, int c0, float pf[8]) const {   // Rectlin mask: the forward activation
    const int n = m / (HC * HC), yx = m % (HC * HC), yy = yx / HC, xx = yx % HC;
    const int y = yy * ST + z / ST, xq = xx * ST + z % ST;
    if (y >= H || xq >= H) { zero8(pf); return; }
    ld8(x + (int64_t(n * H + y) * H + xq) * C + c0, pf);
  }
  __device__ void store8p(int z, int m, int c0, const float v[8], const float xv[8]) const {
    const int n = m / (HC * HC), yx = m % (HC * HC), yy = yx / HC, xx = yx % HC;
    const int y = yy * ST + z / ST, xq = xx * ST + z % ST;
    if (y >= H || xq >= H) return;
    const int64_t i = (int64_t(n * H + y) * H + xq) * C + c0;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = xv[j] > 0.f ? v[j] : 0.f;
    if (dx) st8(dx + i, o);
    if (dx16.hi) umma2::split8_planes(o, dx16.hi + i, dx16.hi + dx16.lo_off + i);
  }
};

// ---- wgrad (MN-major operands, umma_mn.cuh) -----------------------------------------------------
// conv2 / conv3: dW[(r,s,c)][ko] = sum_{n,p,q} X[n, p*ST+r, q*ST+s, c] * dZ[n,p,q,ko]
// One 64-wide m chunk is a contiguous run of the NHWC input: (s, c) are adjacent dims and R*C % 64 == 0.
template <int H, int C, int R, int ST, int KO, int STG = 4>
struct WConvWgrad {
  static constexpr int P = (H - R) / ST + 1, KW = R * R * C;
  static_assert((R * C) % 64 == 0 && KO == 64, "64-element runs must not straddle a filter row");
  static constexpr int kBN = 64, kStages = STG;   // 4 stages = 193 KB (one CTA per SM), 2 stages = 97 KB (two)
  static constexpr bool kAExact = false, kARegs = false, kABulk = false, kFusedUpdate = false;
  PlanePair x16;    // [rows][H][H][C]
  PlanePair dz16;   // [rows][P][P][KO]
  float* part;      // [splits][KW][KO]
  int rows, kb_per_split;
  __device__ int M(int) const { return KW; }
  __device__ int N(int) const { return KO; }
  __device__ void krange(int z, int& kb, int& ke) const {
    const int total = (rows * P * P + 63) / 64;
    kb = min(z * kb_per_split, total);
    ke = min(kb + kb_per_split, total);
  }
  __device__ umma_mn::PixCtx pix(int, int kpix) const {
    const int n = kpix / (P * P), pq = kpix % (P * P);
    return {n, pq / P, pq % P, kpix < rows * P * P};
  }
  __device__ umma2::Planes a_planes(int) const { return {x16.hi, x16.lo_off}; }
  __device__ bool a_run(int, const umma_mn::PixCtx& px, int mchunk, int64_t& off) const {
    const int m = mchunk * 64, r = m / (R * C), sc = m % (R * C);
    off = (int64_t(px.n * H + px.p * ST + r) * H + px.q * ST) * C + sc;
    return m < KW;
  }
  __device__ umma2::Planes b_planes(int) const { return {dz16.hi, dz16.lo_off}; }
  __device__ int64_t b_off(int, const umma_mn::PixCtx& px) const { return (int64_t(px.n * P + px.p) * P + px.q) * KO; }
  __device__ void store8(int z, int m, int n0, const float v[8]) const { st8(part + (int64_t(z) * KW + m) * KO + n0, v); }
};

// conv1: the A operand (row = output pixel, 64 contiguous taps (r,s) of frame c, exact u8 values) is
// exactly the tile conv1_fwd staged for its own MMA, so conv1_fwd ships those tiles to an im2col image
// (V2Conv1Fwd::kDumpA) and this kernel fetches each [64 pixels x 64 taps] sub-tile with ONE TMA bulk copy.
struct WConv1Wgrad {
  static constexpr int kBN = 32, kStages = 4;
  static constexpr bool kAExact = true, kARegs = false, kABulk = true, kFusedUpdate = false;
  const uint8_t* im2col;   // [pixel tile of 128][c = 4][128 x 128 B]
  PlanePair dz16;          // dZ1 [rows][20][20][32]
  float* part;             // [splits][256][32]
  int rows, kb_per_split;
  int tile_rows;           // live rows per 128-row im2col tile: 128 (dense tiling) or 100 (conv1_tma.cuh: 4 tiles/sample)
  __device__ int M(int) const { return kK1; }
  __device__ int N(int) const { return kC1; }
  __device__ int total_kb() const {
    return tile_rows == 128 ? (rows * kP1 * kP1 + 63) / 64 : rows * conv1tma::kTilesPerSample * 2;
  }
  __device__ void krange(int z, int& kb, int& ke) const {
    const int total = total_kb();
    kb = min(z * kb_per_split, total);
    ke = min(kb + kb_per_split, total);
  }
  // kpix indexes the padded im2col rows; n = the real pixel (row of dZ1)
  __device__ umma_mn::PixCtx pix(int, int kpix) const {
    if (tile_rows == 128) return {kpix, 0, 0, kpix < rows * kP1 * kP1};
    const int tile = kpix >> 7, local = kpix & 127;
    return {tile * tile_rows + local, 0, 0, local < tile_rows && tile < rows * conv1tma::kTilesPerSample};
  }
  __device__ const uint8_t* a_sub(int, int c, int kb) const {   // kb = global 64-pixel block
    return im2col + (int64_t(kb >> 1) * (kK1 / 64) + c) * (128 * 128) + (kb & 1) * (64 * 128);
  }
  __device__ umma2::Planes b_planes(int) const { return {dz16.hi, dz16.lo_off}; }
  __device__ int64_t b_off(int, const umma_mn::PixCtx& px) const { return int64_t(px.n) * kC1; }
  __device__ void store8(int z, int m, int n0, const float v[8]) const {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = v[j] * (1.0f / 255.0f);
    st8(part + (int64_t(z) * kK1 + m) * kC1 + n0, o);
  }
};

// fc1: dW4[m][n] = sum_b H3[b][m] * dZ4[b][n]; the reduction rows are the batch samples.
struct WFc1Wgrad {
  static constexpr int kBN = 64, kStages = 2;
  static constexpr bool kAExact = false, kARegs = false, kABulk = false, kFusedUpdate = false;
  PlanePair h3_16;   // [rows][3136]
  PlanePair dz4_16;  // [rows][512]
  float* dw4;        // [3136][512]
  int rows;
  __device__ int M(int) const { return kFlat; }
  __device__ int N(int) const { return kHidden; }
  __device__ void krange(int, int& kb, int& ke) const { kb = 0; ke = (rows + 63) / 64; }
  __device__ umma_mn::PixCtx pix(int, int b) const { return {b, 0, 0, b < rows}; }
  __device__ umma2::Planes a_planes(int) const { return {h3_16.hi, h3_16.lo_off}; }
  __device__ bool a_run(int, const umma_mn::PixCtx& px, int mchunk, int64_t& off) const {
    off = int64_t(px.n) * kFlat + mchunk * 64;
    return mchunk * 64 < kFlat;
  }
  __device__ umma2::Planes b_planes(int) const { return {dz4_16.hi, dz4_16.lo_off}; }
  __device__ int64_t b_off(int, const umma_mn::PixCtx& px) const { return int64_t(px.n) * kHidden; }
  __device__ void store8(int, int m, int n0, const float v[8]) const { st8(dw4 + int64_t(m) * kHidden + n0, v); }
};

// Data-parallel variant: the rows are ALL learners' samples, read from the gather areas that every rank's
// k_xpush fills (comm_p2p.cuh); the parity of the current push epoch selects the area.
struct WFc1WgradGather {
  static constexpr int kBN = 64, kStages = 2;
  static constexpr bool kAExact = false, kARegs = false, kABulk = false, kFusedUpdate = false;
  const __half* h3g;   // [parity][hi | lo][rows][3136]
  const __half* dzg;   // [parity][hi | lo][rows][512]
  int64_t h3_parity, h3_lo, dz_parity, dz_lo;   // elements
  const uint32_t* epoch;   // [0] H3 pushes, [1] dZ4 pushes completed by this rank
  float* dw4;
  int rows;            // world x per-rank minibatch
  __device__ int M(int) const { return kFlat; }
  __device__ int N(int) const { return kHidden; }
  __device__ void krange(int, int& kb, int& ke) const { kb = 0; ke = (rows + 63) / 64; }
  __device__ umma_mn::PixCtx pix(int, int b) const { return {b, 0, 0, b < rows}; }
  __device__ umma2::Planes a_planes(int) const { return {h3g + int64_t(epoch[0] & 1) * h3_parity, h3_lo}; }
  __device__ bool a_run(int, const umma_mn::PixCtx& px, int mchunk, int64_t& off) const {
    off = int64_t(px.n) * kFlat + mchunk * 64;
    return mchunk * 64 < kFlat;
  }
  __device__ umma2::Planes b_planes(int) const { return {dzg + int64_t(epoch[1] & 1) * dz_parity, dz_lo}; }
  __device__ int64_t b_off(int, const umma_mn::PixCtx& px) const { return int64_t(px.n) * kHidden; }
  __device__ void store8(int, int m, int n0, const float v[8]) const { st8(dw4 + int64_t(m) * kHidden + n0, v); }
};

// fc1 wgrad with the optimizer fused into its epilogue (single GPU): the tile of dW4 never leaves the
// SM — RMSProp is applied in place and both fp16 tile images of W4 are refreshed.  Must run after
// fc1_dgrad (which still reads the old dgrad image).
struct WFc1WgradFused {
  static constexpr int kBN = 64, kStages = 2;
  static constexpr bool kAExact = false, kARegs = false, kABulk = false, kFusedUpdate = true;
  PlanePair h3_16;
  PlanePair dz4_16;
  float* w;           // W4 [3136][512]
  float* s;           // RMSProp state
  float* dw_out;      // optional copy of dW4 (b200dqn_net_get_grads), nullptr in production
  uint8_t* img_dgr;   // [25 m-tiles][8 kb][hi 128x128 | lo] — the one fc1 image (forward reads it MN-major)
  int rows;
  OptArgs opt;
  __device__ int M(int) const { return kFlat; }
  __device__ int N(int) const { return kHidden; }
  __device__ void krange(int, int& kb, int& ke) const { kb = 0; ke = (rows + 63) / 64; }
  __device__ umma_mn::PixCtx pix(int, int b) const { return {b, 0, 0, b < rows}; }
  __device__ umma2::Planes a_planes(int) const { return {h3_16.hi, h3_16.lo_off}; }
  __device__ bool a_run(int, const umma_mn::PixCtx& px, int mchunk, int64_t& off) const {
    off = int64_t(px.n) * kFlat + mchunk * 64;
    return mchunk * 64 < kFlat;
  }
  __device__ umma2::Planes b_planes(int) const { return {dz4_16.hi, dz4_16.lo_off}; }
  __device__ int64_t b_off(int, const umma_mn::PixCtx& px) const { return int64_t(px.n) * kHidden; }
  __device__ void store8(int, int, int, const float*) const {}
  __device__ float step_scalar() const { return opt_step_scalar(opt); }
  __device__ void update8(int, int m, int n0, float l, const float g[8], float nw[8]) const {
    const int64_t i = int64_t(m) * kHidden + n0;
    if (dw_out) st8(dw_out + i, g);
    opt_update_vec<8>(opt, l, g, nw, w + i, s + i);   // the configured Neon optimizer (optim.cuh)
    uint4 hi, lo;
    umma::split8(nw, hi, lo);
    uint8_t* base = img_dgr + (int64_t(m / 128) * (kHidden / 64) + n0 / 64) * (128 * 256) +
                    umma::sw128_off(m % 128, (n0 % 64) / 8);
    *reinterpret_cast<uint4*>(base) = hi;
    *reinterpret_cast<uint4*>(base + 128 * 128) = lo;
  }
  __device__ void pack_col8(int, int, int, const float*) const {}   // no column-oriented image any more
};

// ---- weight tile-image sources (k_pack_image) --------------------------------------------------
template <int K, int N>
struct PackFwdConv {   // B operand of a forward conv: rows = output channel n, K = filter taps
  static constexpr bool kRowMajorThreads = false;
  const float* w;      // [K][N]
  __host__ __device__ int tiles() const { return 1; }
  __host__ __device__ int rows() const { return N; }
  __host__ __device__ int kblocks() const { return K / 64; }
  __device__ void src8(int, int r, int k0, float v[8]) const {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = w[(k0 + j) * N + r];
  }
};
struct PackFc1Dgrad {  // the fc1 image: rows = flat index m, 64 hidden units per row (dgrad: K-major A; forward: MN-major A)
  static constexpr bool kRowMajorThreads = true;
  const float* w;
  __host__ __device__ int tiles() const { return (kFlat + 127) / 128; }
  __host__ __device__ int rows() const { return 128; }
  __host__ __device__ int kblocks() const { return kHidden / 64; }
  __device__ void src8(int tile, int r, int k0, float v[8]) const {
    const int m = tile * 128 + r;
    if (m >= kFlat) { zero8(v); return; }
    ld8(w + m * kHidden + k0, v);
  }
};
template <int H, int C, int R, int ST, int KO>
struct PackConvDgrad { // B operand of a conv dgrad, one tile per output-parity class
  static constexpr int RT = R / ST, K = RT * RT * KO;
  static constexpr bool kRowMajorThreads = true;
  const float* w;      // [(r,s,c)][KO]
  __host__ __device__ int tiles() const { return ST * ST; }
  __host__ __device__ int rows() const { return C; }
  __host__ __device__ int kblocks() const { return K / 64; }
  __device__ void src8(int z, int c, int k0, float v[8]) const {
    const int rp = k0 / (RT * KO), sp = (k0 / KO) % RT, ko = k0 % KO;
    const int r = rp * ST + z / ST, s = sp * ST + z % ST;
    ld8(w + ((r * R + s) * C + c) * KO + ko, v);
  }
};

// ------------------------------------------------------------------------------------------
// Conv-layer optimizer, fused: split-K partial reduction (8 lanes per element, fixed tree ->
// deterministic) + Neon RMSProp (same operation order as k_optimizer, bit-exact given equal gradients)
// + refresh of the layer's fp16 tile images (forward B operand; dgrad B operand for conv2/conv3).
// Replaces three launches (optimizer, pack fwd, pack dgrad) on the tail of the step.
// ------------------------------------------------------------------------------------------
// XCHG (data-parallel learners, experimental — see umma_opt_conv_xll): between the reduction and the update the
// 8 lanes of an element exchange it with the other ranks in the LL protocol of comm_p2p.cuh — lane p pushes this
// rank's float4 to rank p and polls rank p's line — and the W contributions are added in rank order.
template <int KR, int N, bool DGRAD, int C, int R, int ST, bool XCHG = false>
__global__ void __launch_bounds__(256)
k_opt_conv(const float* __restrict__ part, int splits, float* __restrict__ w, float* __restrict__ sst,
           uint8_t* __restrict__ img_fwd, uint8_t* __restrict__ img_dgr, const OptArgs opt, const XllArgs x,
           const KTrace kt) {
  constexpr int64_t kSize = int64_t(KR) * N;
  kt_begin(kt);
  pdl_wait();
  pdl_launch_dependents();
  const int tid = threadIdx.x, lane8 = tid & 7;
  uint32_t epoch = 0;
  if constexpr (XCHG) epoch = *reinterpret_cast<volatile uint32_t*>(x.epoch + x.chan) + 1;
  const int64_t e4 = blockIdx.x * 32 + (tid >> 3);
  const int64_t i = e4 * 4;
  const bool live = i < kSize;
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) {
    const float* p = part + i;
    float4 v[8];
    int cnt = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) {          // up to 64 splits: 8 independent loads per lane
      const int sp = lane8 + 8 * u;
      if (sp < splits) { v[u] = *reinterpret_cast<const float4*>(p + sp * kSize); cnt = u + 1; }
    }
    for (int u = 0; u < cnt; ++u) { g.x += v[u].x; g.y += v[u].y; g.z += v[u].z; g.w += v[u].w; }
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    g.x += __shfl_xor_sync(0xffffffffu, g.x, o);
    g.y += __shfl_xor_sync(0xffffffffu, g.y, o);
    g.z += __shfl_xor_sync(0xffffffffu, g.z, o);
    g.w += __shfl_xor_sync(0xffffffffu, g.w, o);
  }
  if constexpr (XCHG) {
    float4 v = lane8 == x.rank ? g : make_float4(0.f, 0.f, 0.f, 0.f);
    if (live && lane8 < x.world && lane8 != x.rank) {
      const int64_t par = int64_t(epoch & 1) * x.world * x.lines_per_src, el = (x.ll4 + e4) * 2;
      uint4* dst = x.recv[lane8] + par + int64_t(x.rank) * x.lines_per_src + el;
      st_ll(dst, __float_as_uint(g.x), __float_as_uint(g.y), epoch);
      st_ll(dst + 1, __float_as_uint(g.z), __float_as_uint(g.w), epoch);
      const uint4* src = x.recv[x.rank] + par + int64_t(lane8) * x.lines_per_src + el;
      ll_wait(src, epoch, x.err, v.x, v.y);
      ll_wait(src + 1, epoch, x.err, v.z, v.w);
    }
    __syncwarp();
    const int base = (tid & 31) & ~7;
    g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < x.world; ++p) {   // rank order, the same on every rank
      const float a = __shfl_sync(0xffffffffu, v.x, base + p), b = __shfl_sync(0xffffffffu, v.y, base + p);
      const float c = __shfl_sync(0xffffffffu, v.z, base + p), d = __shfl_sync(0xffffffffu, v.w, base + p);
      if (p == 0) g = make_float4(a, b, c, d);
      else { g.x += a; g.y += b; g.z += c; g.w += d; }
    }
  }
  if (live && lane8 == 0) {
    float wp[4];
    opt_update_vec<4>(opt, opt_step_scalar(opt), reinterpret_cast<const float*>(&g), wp, w + i, sst + i);
    __half hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) umma2::split1(wp[j], hi[j], lo[j]);
    const int k = int(i / N), n0 = int(i % N);
    {  // forward image: rows = output channel n, 8-element chunks along k
      uint8_t* base = img_fwd + int64_t(k / 64) * (N * 256) + (k % 8) * 2;
      const int c8 = (k % 64) / 8;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t off = umma::sw128_off(n0 + j, c8);
        *reinterpret_cast<__half*>(base + off) = hi[j];
        *reinterpret_cast<__half*>(base + N * 128 + off) = lo[j];
      }
    }
    if constexpr (DGRAD) {  // dgrad image: one tile per output-parity class, rows = input channel c, K = (r', s', ko)
      constexpr int RT = R / ST, NKB = RT * RT * N / 64;
      const int r = k / (R * C), s = (k / C) % R, c = k % C;
      const int z = (r % ST) * ST + (s % ST);
      const int kd = ((r / ST) * RT + (s / ST)) * N + n0;
      uint8_t* base = img_dgr + (int64_t(z) * NKB + kd / 64) * (C * 256) + umma::sw128_off(c, (kd % 64) / 8) + (kd % 8) * 2;
      *reinterpret_cast<uint2*>(base) = *reinterpret_cast<const uint2*>(hi);
      *reinterpret_cast<uint2*>(base + C * 128) = *reinterpret_cast<const uint2*>(lo);
    }
  }
  if constexpr (XCHG) {   // the last block to finish publishes the layer's new epoch (as in k_xll)
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      if (atomicAdd(x.ticket + x.chan, 1u) == gridDim.x - 1) {
        x.ticket[x.chan] = 0;
        __threadfence();
        *reinterpret_cast<volatile uint32_t*>(x.epoch + x.chan) = epoch;
      }
    }
  }
  kt_end(kt);
}

// fc1 optimizer: RMSProp on 8 consecutive hidden units of one flat index per thread — which is exactly
// one 16-byte chunk of the row-oriented (dgrad) tile image, refreshed in the same pass; the
// column-oriented (forward) image is rebuilt by k_pack_image right after.  Both kernels are smem-free,
// light on registers and launched on a CAPPED grid (2 CTAs per SM, grid-stride loop) so that they
// co-reside with the tcgen05 kernels of the critical chain instead of locking them out of the SMs.
__global__ void __launch_bounds__(256)
k_opt_fc1(const float* __restrict__ dw, float* __restrict__ w, float* __restrict__ sst, uint8_t* __restrict__ img_dgr,
          const OptArgs opt, const uint32_t* __restrict__ gate, const KTrace kt) {
  kt_begin(kt);
  pdl_wait();
  pdl_launch_dependents();
  if (gate && *gate == 0) {   // nothing pending (first step after a flush): uniform across the grid
    kt_end(kt);
    return;
  }
  constexpr int kNB = kHidden / 8;
  const float l_step = opt_step_scalar(opt);
  for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < kFlat * kNB; id += gridDim.x * blockDim.x) {
    const int m = id / kNB, n0 = (id % kNB) * 8;
    const int64_t i = int64_t(m) * kHidden + n0;
    float g[8], wv[8];
    ld8(dw + i, g);
    opt_update_vec<8>(opt, l_step, g, wv, w + i, sst + i);   // the configured Neon optimizer (optim.cuh)
    uint4 hi, lo;
    umma::split8(wv, hi, lo);
    uint8_t* base = img_dgr + (int64_t(m / 128) * (kHidden / 64) + n0 / 64) * (128 * 256) +
                    umma::sw128_off(m % 128, (n0 % 64) / 8);
    *reinterpret_cast<uint4*>(base) = hi;
    *reinterpret_cast<uint4*>(base + 128 * 128) = lo;
  }
  kt_end(kt);
}

int umma_opt_fc1(b200dqn_net* n, int rows, cudaStream_t st, bool from_g, const uint32_t* gate) {
  UmmaState* u = ust(n);
  const LayerTable& lt = n->lt;
  const float* dw = from_g ? n->d_g + lt.off[3] : n->d_part + lt.part_off[3];
  // One kernel, one image: the update refreshes the row-oriented tile image in the same pass, and the forward reads
  // that image too (MN-major) — the column-oriented forward image and its re-pack kernel (round 1-2: 6.4 MB read +
  // 6.4 MB written per step, 5 us at the end of the fc1 branch) are gone.
  // capped grid (CTAs per SM, grid-stride): the kernel shares the SMs — and the L2 — with the dgrad chain
  static const int per_sm = getenv("B200DQN_OPT_FC1_CTAS") ? atoi(getenv("B200DQN_OPT_FC1_CTAS")) : 2;
  const int ctas = per_sm > 0 ? per_sm * n->sm_count : n->sm_count / (-per_sm > 0 ? -per_sm : 1);
  B2_CHECK_CUDA(launch_pdl(k_opt_fc1, dim3(ctas), dim3(256), 0, st, dw, n->d_w + lt.off[3],
                           n->d_s + lt.off[3], u->img_dgr[0], make_opt_args(n, rows), gate, ktrace_slot("opt_fc1")));
  B2_PROF("opt_fc1", st);
  return B200DQN_OK;
}

// RMSProp + image refresh of conv layer l (0..2), fused (single-GPU path of the tcgen05 engine).
int umma_opt_conv(b200dqn_net* n, int l, int rows, cudaStream_t st, const char* label, bool from_g) {
  UmmaState* u = ust(n);
  const LayerTable& lt = n->lt;
  const OptArgs opt = make_opt_args(n, rows);
  const float* part = from_g ? n->d_g + lt.off[l] : n->d_part + lt.part_off[l];
  const int nsplits = from_g ? 1 : lt.splits[l];
  float* w = n->d_w + lt.off[l];
  float* s = n->d_s + lt.off[l];
  B2_REQUIRE(nsplits <= 64, B200DQN_EINVAL, "k_opt_conv handles at most 64 split-K partials");
  const int64_t size = lt.off[l + 1] - lt.off[l];
  const dim3 grid(unsigned((size / 4 + 31) / 32)), block(256);
  cudaError_t e;
  const XllArgs none{};
  if (l == 0)
    e = launch_pdl(k_opt_conv<kK1, kC1, false, 4, 8, 4>, grid, block, 0, st, part, nsplits, w, s, u->img_fwd[0][0],
                   (uint8_t*)nullptr, opt, none, ktrace_slot(label));
  else if (l == 1)
    e = launch_pdl(k_opt_conv<kK2, kC2, true, kC1, 4, 2>, grid, block, 0, st, part, nsplits, w, s, u->img_fwd[0][1],
                   u->img_dgr[2], opt, none, ktrace_slot(label));
  else
    e = launch_pdl(k_opt_conv<kK3, kC3, true, kC2, 3, 1>, grid, block, 0, st, part, nsplits, w, s, u->img_fwd[0][2],
                   u->img_dgr[1], opt, none, ktrace_slot(label));
  B2_CHECK_CUDA(e);
  B2_PROF(label, st);
  return B200DQN_OK;
}

// EXPERIMENTAL (B200DQN_FUSED_XLL=1; written after round 1's GPU budget was spent, not yet run on hardware):
// split-K reduction + LL all-reduce across the learners + RMSProp + image refresh of conv layer l in ONE launch
// instead of three (reduce, k_xll, k_opt_conv) on the tail of the data-parallel step.
int umma_opt_conv_xll(b200dqn_net* n, int l, int rows, cudaStream_t st, const char* label) {
  UmmaState* u = ust(n);
  const LayerTable& lt = n->lt;
  const OptArgs opt = make_opt_args(n, rows);
  B2_REQUIRE(l >= 0 && l < 3 && lt.splits[l] <= 64 && n->world <= 8, B200DQN_EINVAL, "opt_conv_xll: bad layer / world");
  XllArgs x{};
  int rc = comm_xll_args(n, l, &x);
  if (rc) return rc;
  const float* part = n->d_part + lt.part_off[l];
  float* w = n->d_w + lt.off[l];
  float* s = n->d_s + lt.off[l];
  const int64_t size = lt.off[l + 1] - lt.off[l];
  const dim3 grid(unsigned((size / 4 + 31) / 32)), block(256);
  cudaError_t e;
  if (l == 0)
    e = launch_pdl(k_opt_conv<kK1, kC1, false, 4, 8, 4, true>, grid, block, 0, st, part, lt.splits[l], w, s,
                   u->img_fwd[0][0], (uint8_t*)nullptr, opt, x, ktrace_slot(label));
  else if (l == 1)
    e = launch_pdl(k_opt_conv<kK2, kC2, true, kC1, 4, 2, true>, grid, block, 0, st, part, lt.splits[l], w, s,
                   u->img_fwd[0][1], u->img_dgr[2], opt, x, ktrace_slot(label));
  else
    e = launch_pdl(k_opt_conv<kK3, kC3, true, kC2, 3, 1, true>, grid, block, 0, st, part, lt.splits[l], w, s,
                   u->img_fwd[0][2], u->img_dgr[1], opt, x, ktrace_slot(label));
  B2_CHECK_CUDA(e);
  B2_PROF(label, st);
  return B200DQN_OK;
}

static int64_t fwd_image_bytes(int layer) {
  switch (layer) {
    case 0: return int64_t(kK1 / 64) * kC1 * 256;
    case 1: return int64_t(kK2 / 64) * kC2 * 256;
    case 2: return int64_t(kK3 / 64) * kC3 * 256;
    default: return int64_t((kFlat + 127) / 128) * (kHidden / 64) * 128 * 256;   // fc1: the row-oriented image
  }
}

// (re)build the tile images of layers [l0, l1] of network `which` from its fp32 master weights
int umma_pack_layers(b200dqn_net* n, int which, int l0, int l1, cudaStream_t st) {
  if (n->cfg.math_mode != B200DQN_MATH_TCGEN05) return B200DQN_OK;
  UmmaState* u = ust(n);
  const LayerTable& lt = n->lt;
  const float* w = which ? n->d_tw : n->d_w;
  int rc = 0;
  for (int l = l0; l <= l1 && !rc; ++l) {
    switch (l) {
      case 0: rc = umma2::launch_pack("pack_c1", PackFwdConv<kK1, kC1>{w + lt.off[0]}, u->img_fwd[which][0], st); break;
      case 1:
        rc = umma2::launch_pack("pack_c2f", PackFwdConv<kK2, kC2>{w + lt.off[1]}, u->img_fwd[which][1], st);
        if (!rc && !which)
          rc = umma2::launch_pack("pack_c2d", PackConvDgrad<kP1, kC1, 4, 2, kC2>{w + lt.off[1]}, u->img_dgr[2], st);
        break;
      case 2:
        rc = umma2::launch_pack("pack_c3f", PackFwdConv<kK3, kC3>{w + lt.off[2]}, u->img_fwd[which][2], st);
        if (!rc && !which)
          rc = umma2::launch_pack("pack_c3d", PackConvDgrad<kP2, kC2, 3, 1, kC3>{w + lt.off[2]}, u->img_dgr[1], st);
        break;
      case 3:
        rc = umma2::launch_pack("pack_fc1", PackFc1Dgrad{w + lt.off[3]}, u->img_fwd[which][3], st);
        break;
      default: break;  // fc2 runs on CUDA cores (N = A <= 18)
    }
  }
  return rc;
}

// fc1 forward split-K over blockIdx.z (the head kernel sums the partials): 49 k-blocks of 64 -> 7 per CTA,
// 4 M-tiles x 7 x 2 nets = 56 CTAs at batch 32; large minibatches bring their own tiles, so fewer splits keep the
// partial-sum traffic down.  (13 splits = 104 CTAs was measured: fc1_fwd 6.3 -> 12 us inside the step — see below.)
static inline int fc1_splits_for(int rows) {
  static const int forced = getenv("B200DQN_FC1_SPLITS") ? atoi(getenv("B200DQN_FC1_SPLITS")) : 0;
  if (forced >= 1 && forced <= kFc1Splits) return forced;
  return rows <= 256 ? 7 : 4;
}
// Cluster split-K (umma2.cuh) for the kernels whose tile count leaves most of the 148 SMs idle at batch 32:
//   conv2_fwd 42 tiles x 3 partners (8 k-blocks -> 3/3/2),  conv3_fwd 26 x 4 (9 -> 3/2/2/2),
//   conv3_dgrad 21 x 4,  fc1_dgrad 25 x 4 (8 -> 2 each).
// OFF by default (B200DQN_SPLITK=1 turns it on): parity-clean, but MEASURED SLOWER inside the step — 85.2 us vs
// 73.6 us per step on a B200 (profiles/r2c_*): with 100+ CTAs of 193 KB shared memory per kernel the successor of the
// PDL chain finds no free SM to pre-launch on (its prologue no longer hides behind the predecessor) and the cluster
// needs all its SMs at once; the few-CTA tiles win because consecutive kernels CO-RESIDE.
static const bool g_splitk = getenv("B200DQN_SPLITK") && atoi(getenv("B200DQN_SPLITK")) != 0;
// Larger minibatches already fill the chip with tiles: split only while the tile count is below the SM count.
static inline bool use_splitk(int tiles, int ks) { return g_splitk && tiles * ks <= 160; }
constexpr int kUWgradKb = 4;      // minimum k-blocks (of 64 pixels) per wgrad split

// k-blocks (of 64 pixels) per wgrad split: at least kUWgradKb, and few enough splits (<= 48) for the
// one-pass reduction of k_opt_conv
// conv1 through tensor-map TMA (conv1_tma.cuh) unless B200DQN_CONV1=ldg selects the register-path gather of umma2.cuh
static const bool g_conv1_tma = !(getenv("B200DQN_CONV1") && strcmp(getenv("B200DQN_CONV1"), "ldg") == 0);
static inline int conv1_pixels_padded(int rows) { return g_conv1_tma ? rows * conv1tma::kTilesPerSample * 128 : rows * kP1 * kP1; }

int umma_wgrad_kb(int layer, int rows) {
  const int kred = layer == 0 ? conv1_pixels_padded(rows) : layer == 1 ? rows * kP2 * kP2 : rows * kP3 * kP3;
  const int kbs = (kred + 63) / 64;
  int per = (kbs + 47) / 48;
  if (layer == 0 && per < 8) per = 8;   // conv1: A tiles arrive by TMA bulk copies; 8 k-blocks per CTA keep the split count (and
                                        // the optimizer's partial-sum loads) down: measured 43 splits -> opt_conv1 +1.5 us
  return per > kUWgradKb ? per : kUWgradKb;
}
int umma_wgrad_splits(int layer, int rows) {
  const int kred = layer == 0 ? conv1_pixels_padded(rows) : layer == 1 ? rows * kP2 * kP2 : rows * kP3 * kP3;
  const int kbs = (kred + 63) / 64, per = umma_wgrad_kb(layer, rows);
  return (kbs + per - 1) / per;
}

int umma_net_init(b200dqn_net* n) {
  if (n->cfg.math_mode != B200DQN_MATH_TCGEN05) return B200DQN_OK;
  auto* u = new UmmaState();
  n->umma_state = u;
  const int nb = n->nb;
  u->h_elems[0] = int64_t(nb) * kP1 * kP1 * kC1;
  u->h_elems[1] = int64_t(nb) * kP2 * kP2 * kC2;
  u->h_elems[2] = int64_t(nb) * kFlat;
  u->dz_elems[0] = int64_t(nb) * kHidden;
  u->dz_elems[1] = int64_t(nb) * kFlat;
  u->dz_elems[2] = int64_t(nb) * kP2 * kP2 * kC2;
  u->dz_elems[3] = int64_t(nb) * kP1 * kP1 * kC1;
  for (int i = 0; i < 3; ++i) {
    for (int z = 0; z < 2; ++z) {
      B2_CHECK_CUDA(cudaMalloc(&u->h16[i][z], 2 * u->h_elems[i] * sizeof(__half)));
      B2_CHECK_CUDA(cudaMemset(u->h16[i][z], 0, 2 * u->h_elems[i] * sizeof(__half)));
    }
  }
  for (int i = 0; i < 4; ++i) {
    B2_CHECK_CUDA(cudaMalloc(&u->dz16[i], 2 * u->dz_elems[i] * sizeof(__half)));
    B2_CHECK_CUDA(cudaMemset(u->dz16[i], 0, 2 * u->dz_elems[i] * sizeof(__half)));
  }
  for (int l = 0; l < 4; ++l) {
    u->img_fwd_bytes[l] = fwd_image_bytes(l);
    for (int z = 0; z < 2; ++z) {
      if (z == 1 && n->d_tw == n->d_w) { u->img_fwd[1][l] = u->img_fwd[0][l]; continue; }
      B2_CHECK_CUDA(cudaMalloc(&u->img_fwd[z][l], u->img_fwd_bytes[l]));
      B2_CHECK_CUDA(cudaMemset(u->img_fwd[z][l], 0, u->img_fwd_bytes[l]));
    }
  }
  B2_CHECK_CUDA(cudaMalloc(&u->im2col1, int64_t(nb) * conv1tma::kTilesPerSample * (kK1 / 64) * 128 * 128));
  u->img_dgr[0] = u->img_fwd[0][3];   // fc1: ONE row-oriented image serves the dgrad (K-major) and the forward (MN-major)
  const int64_t dgr_bytes[3] = {0, int64_t(kK3 / 64) * kC2 * 256, int64_t(4) * (256 / 64) * kC1 * 256};
  for (int i = 1; i < 3; ++i) {
    B2_CHECK_CUDA(cudaMalloc(&u->img_dgr[i], dgr_bytes[i]));
    B2_CHECK_CUDA(cudaMemset(u->img_dgr[i], 0, dgr_bytes[i]));
  }
  return B200DQN_OK;
}

void umma_net_destroy(b200dqn_net* n) {
  UmmaState* u = ust(n);
  if (!u) return;
  for (int i = 0; i < 3; ++i) {
    for (int z = 0; z < 2; ++z) cudaFree(u->h16[i][z]);
    if (i > 0) cudaFree(u->img_dgr[i]);   // [0] aliases img_fwd[0][3]
  }
  for (int i = 0; i < 4; ++i) cudaFree(u->dz16[i]);
  cudaFree(u->im2col1);
  for (int l = 0; l < 4; ++l) {
    if (u->img_fwd[1][l] != u->img_fwd[0][l]) cudaFree(u->img_fwd[1][l]);
    cudaFree(u->img_fwd[0][l]);
  }
  delete u;
  n->umma_state = nullptr;
}

int umma_weights_changed(b200dqn_net* n, cudaStream_t st) {
  if (n->cfg.math_mode != B200DQN_MATH_TCGEN05) return B200DQN_OK;
  int rc = umma_pack_layers(n, 0, 0, 3, st);
  if (!rc && n->d_tw != n->d_w) rc = umma_pack_layers(n, 1, 0, 3, st);
  return rc;
}

int umma_target_synced(b200dqn_net* n, cudaStream_t st) {
  if (n->cfg.math_mode != B200DQN_MATH_TCGEN05 || n->d_tw == n->d_w) return B200DQN_OK;
  UmmaState* u = ust(n);
  for (int l = 0; l < 4; ++l)
    B2_CHECK_CUDA(cudaMemcpyAsync(u->img_fwd[1][l], u->img_fwd[0][l], u->img_fwd_bytes[l], cudaMemcpyDeviceToDevice, st));
  return B200DQN_OK;
}

void umma_dz4_planes(b200dqn_net* n, __half** hi, int64_t* lo_off) {
  UmmaState* u = ust(n);
  *hi = u ? u->dz16[0] : nullptr;
  *lo_off = u ? u->dz_elems[0] : 0;
}

int umma_forward(b200dqn_net* n, const uint8_t* const src[2], const int32_t* const idx[2], const int shift[2],
                 const int64_t nframes[2], int nets, int rows, cudaStream_t st) {
  UmmaState* u = ust(n);
  int rc;
  auto planes = [&](int i, int z) { return PlanePair{u->h16[i][z], u->h_elems[i]}; };
  if (g_conv1_tma) {
    // frames by tensor-map TMA.  Ring case (both states out of one frame array, poststates one frame later): ONE
    // 5-frame window per CTA serves both networks.
    conv1tma::Params p{};
    const bool shared5 = nets == 2 && src[0] == src[1] && idx[0] == idx[1] && shift[1] == shift[0] + 1;
    for (int z = 0; z < 2; ++z) {
      p.idx[z] = idx[z]; p.shift[z] = shift[z];
      p.wimg[z] = u->img_fwd[z][0]; p.out16[z] = u->h16[0][z];
    }
    p.out[0] = n->d_h1[0];
    p.out[1] = nullptr;                       // nothing reads the target network's fp32 H1
    p.shared5 = shared5 ? 1 : 0; p.nets = nets; p.rows = rows; p.lo_off = u->h_elems[0];
    p.im2col = (nets == 2 && rows == n->nb) ? u->im2col1 : nullptr;
    CUtensorMap m0, m1;
    const int64_t nframes0 = nframes[0], nframes1 = nframes[1];
    if ((rc = conv1tma::make_frame_map(&m0, src[0], nframes0, shared5 ? kHist + 1 : kHist))) return rc;
    if ((rc = conv1tma::make_frame_map(&m1, src[nets == 2 ? 1 : 0], nets == 2 ? nframes1 : nframes0, kHist))) return rc;
    static bool configured = false;
    if (!configured) {
      B2_CHECK_CUDA(cudaFuncSetAttribute(conv1tma::k_conv1_tma, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         int(conv1tma::smem_bytes(2))));
      configured = true;
    }
    const uint32_t smem = conv1tma::smem_bytes((nets == 2 && !shared5) ? 2 : 1);
    B2_CHECK_CUDA(launch_pdl(conv1tma::k_conv1_tma, dim3(rows * conv1tma::kTilesPerSample), dim3(umma2::kThreads2),
                             smem, st, m0, m1, p, ktrace_slot("conv1_fwd")));
    B2_PROF("conv1_fwd", st);
  } else {
    V2Conv1Fwd p;
    for (int z = 0; z < 2; ++z) {
      p.src[z] = src[z]; p.idx[z] = idx[z]; p.shift[z] = shift[z];
      p.wimg[z] = u->img_fwd[z][0]; p.out[z] = n->d_h1[z]; p.out16[z] = planes(0, z);
    }
    p.rows = rows;
    p.im2col = (nets == 2 && rows == n->nb) ? u->im2col1 : nullptr;   // only a train step feeds conv1_wgrad
    if ((rc = umma2::launch_umma2("conv1_fwd", p, rows * kP1 * kP1, kC1, nets, st))) return rc;
  }
  {
    using P = V2ConvFwd<kP1, kC1, 4, 2, kC2>;
    P p;
    for (int z = 0; z < 2; ++z) {
      p.in16[z] = planes(0, z); p.wimg[z] = u->img_fwd[z][1]; p.out16[z] = planes(1, z);
      p.out[z] = z ? nullptr : n->d_h2[z];     // nothing reads the target network's fp32 activations
    }
    p.rows = rows;
    const int tiles = (rows * kP2 * kP2 + 127) / 128 * nets;
    if (use_splitk(tiles, 3)) rc = umma2::launch_umma2<P, 3>("conv2_fwd", p, rows * kP2 * kP2, kC2, nets, st);
    else rc = umma2::launch_umma2("conv2_fwd", p, rows * kP2 * kP2, kC2, nets, st);
    if (rc) return rc;
  }
  {
    using P = V2ConvFwd<kP2, kC2, 3, 1, kC3>;
    P p;
    for (int z = 0; z < 2; ++z) {
      p.in16[z] = planes(1, z); p.wimg[z] = u->img_fwd[z][2]; p.out16[z] = planes(2, z);
      p.out[z] = z ? nullptr : n->d_h3[z];
    }
    p.rows = rows;
    const int tiles = (rows * kP3 * kP3 + 127) / 128 * nets;
    if (use_splitk(tiles, 4)) rc = umma2::launch_umma2<P, 4>("conv3_fwd", p, rows * kP3 * kP3, kC3, nets, st);
    else rc = umma2::launch_umma2("conv3_fwd", p, rows * kP3 * kP3, kC3, nets, st);
    if (rc) return rc;
    // data-parallel learners: this rank's H3 rows start travelling to every rank's fc1_wgrad now
    if (nets == 2 && rows == n->nb && comm_gather_active(n, st) && (rc = umma_push_h3(n, st))) return rc;
  }
  {
    V2Fc1Fwd p;
    for (int z = 0; z < 2; ++z) { p.in16[z] = planes(2, z); p.wimg[z] = u->img_fwd[z][3]; }
    p.part = n->d_fc1part; p.rows = rows; p.splits = fc1_splits_for(rows);
    if (n->defer_fc1) {
      // the previous step's fc1 update (side branch since the start of this step) must have refreshed the image; the
      // kernel then has two parents, so it is launched as an ordinary node (its pre-wait weight prefetch would
      // otherwise run ahead of the join)
      static const bool keep_pdl = getenv("B200DQN_DEFER_PDL") != nullptr;
      B2_CHECK_CUDA(cudaStreamWaitEvent(st, n->ev[16], 0));
      if (keep_pdl) {
        rc = umma2::launch_umma2("fc1_fwd", p, kHidden, rows, nets * p.splits, st);
      } else {
        NoPdlScope plain;
        rc = umma2::launch_umma2("fc1_fwd", p, kHidden, rows, nets * p.splits, st);
      }
      if (rc) return rc;
    } else if ((rc = umma2::launch_umma2("fc1_fwd", p, kHidden, rows, nets * p.splits, st))) return rc;
  }
  return B200DQN_OK;
}

// B200DQN_STAGES2=label,label,...: run that kernel with a 2-stage operand ring (about half the shared memory, so two
// CTAs of the step's kernels fit on an SM) instead of the deepest ring.
// Default on ONE GPU: conv2_dgrad — 100 CTAs of 4 k-blocks each; at 81 KB instead of 161 KB they occupy 50 SMs instead
// of 100 while conv3_wgrad / conv2_wgrad / conv1_wgrad look for SMs (measured period 73.4 -> 71.9 us,
// profiles/r2o_periods.txt; the conv wgrads themselves are slower with the shallow ring).  With data-parallel learners
// the same choice costs 12 us per step (2 x B200: 96.0 vs 83.6 us, profiles/r2s_*), so there the default is none.
static bool shallow_ring(const b200dqn_net* net, const char* label) {
  static const char* env = getenv("B200DQN_STAGES2");
  const char* list = env ? env : net->world == 1 ? "conv2_dgrad" : "";
  const size_t n = strlen(label);
  for (const char* p = list; (p = strstr(p, label)) != nullptr; p += n)
    if ((p == list || p[-1] == ',') && (p[n] == 0 || p[n] == ',')) return true;
  return false;
}

int umma_backward_op(b200dqn_net* n, int op, const uint8_t* src, const int32_t* idx, int shift, int rows,
                     cudaStream_t st) {
  const LayerTable& lt = n->lt;
  const float* w = n->d_w;
  switch (op) {
    case 0: {
      UmmaState* u = ust(n);
      WFc1Wgrad p{PlanePair{u->h16[2][0], u->h_elems[2]}, PlanePair{u->dz16[0], u->dz_elems[0]},
                  n->d_part + lt.part_off[3], rows};
      return umma_mn::launch_umma_mn("fc1_wgrad", p, kFlat, kHidden, 1, st);
    }
    case 1: {
      UmmaState* u = ust(n);
      // the fp32 copies of dZ3/dZ2/dZ1 have no reader in this engine (wgrads and dgrads take the fp16 planes)
      V2Fc1Dgrad p{u->img_dgr[0], PlanePair{u->dz16[0], u->dz_elems[0]}, n->d_h3[0], n->keep_grads ? n->d_dz3 : nullptr,
                   PlanePair{u->dz16[1], u->dz_elems[1]}, rows};
      if (use_splitk(25 * ((rows + 31) / 32), 4)) return umma2::launch_umma2<V2Fc1Dgrad, 4>("fc1_dgrad", p, kFlat, rows, 1, st);
      return umma2::launch_umma2("fc1_dgrad", p, kFlat, rows, 1, st);
    }
    case 2: {
      UmmaState* u = ust(n);
      using P = WConvWgrad<kP2, kC2, 3, 1, kC3>;
      using P2 = WConvWgrad<kP2, kC2, 3, 1, kC3, 2>;
      if (shallow_ring(n, "conv3_wgrad")) {
        P2 p{PlanePair{u->h16[1][0], u->h_elems[1]}, PlanePair{u->dz16[1], u->dz_elems[1]},
             n->d_part + lt.part_off[2], rows, umma_wgrad_kb(2, rows)};
        return umma_mn::launch_umma_mn("conv3_wgrad", p, P::KW, kC3, lt.splits[2], st);
      }
      P p{PlanePair{u->h16[1][0], u->h_elems[1]}, PlanePair{u->dz16[1], u->dz_elems[1]},
          n->d_part + lt.part_off[2], rows, umma_wgrad_kb(2, rows)};
      return umma_mn::launch_umma_mn("conv3_wgrad", p, P::KW, kC3, lt.splits[2], st);
    }
    case 3: {
      UmmaState* u = ust(n);
      using P = V2ConvDgrad<kP2, kC2, 3, 1, kC3>;
      P p{PlanePair{u->dz16[1], u->dz_elems[1]}, u->img_dgr[1], n->d_h2[0], n->keep_grads ? n->d_dz2 : nullptr,
          PlanePair{u->dz16[2], u->dz_elems[2]}, rows};
      if (use_splitk((rows * P::HC * P::HC + 127) / 128, 4))
        return umma2::launch_umma2<P, 4>("conv3_dgrad", p, rows * P::HC * P::HC, kC2, 1, st);
      return umma2::launch_umma2("conv3_dgrad", p, rows * P::HC * P::HC, kC2, 1, st);
    }
    case 4: {
      UmmaState* u = ust(n);
      using P = WConvWgrad<kP1, kC1, 4, 2, kC2>;
      using P2 = WConvWgrad<kP1, kC1, 4, 2, kC2, 2>;
      if (shallow_ring(n, "conv2_wgrad")) {
        P2 p{PlanePair{u->h16[0][0], u->h_elems[0]}, PlanePair{u->dz16[2], u->dz_elems[2]},
             n->d_part + lt.part_off[1], rows, umma_wgrad_kb(1, rows)};
        return umma_mn::launch_umma_mn("conv2_wgrad", p, P::KW, kC2, lt.splits[1], st);
      }
      P p{PlanePair{u->h16[0][0], u->h_elems[0]}, PlanePair{u->dz16[2], u->dz_elems[2]},
          n->d_part + lt.part_off[1], rows, umma_wgrad_kb(1, rows)};
      return umma_mn::launch_umma_mn("conv2_wgrad", p, P::KW, kC2, lt.splits[1], st);
    }
    case 5: {
      UmmaState* u = ust(n);
      using P = V2ConvDgrad<kP1, kC1, 4, 2, kC2>;
      using P2 = V2ConvDgrad<kP1, kC1, 4, 2, kC2, 2>;
      if (shallow_ring(n, "conv2_dgrad")) {   // 81 KB per CTA: the 100 CTAs take 50 SMs instead of 100
        P2 p{PlanePair{u->dz16[2], u->dz_elems[2]}, u->img_dgr[2], n->d_h1[0], n->keep_grads ? n->d_dz1 : nullptr,
             PlanePair{u->dz16[3], u->dz_elems[3]}, rows};
        return umma2::launch_umma2("conv2_dgrad", p, rows * P::HC * P::HC, kC1, 4, st);
      }
      P p{PlanePair{u->dz16[2], u->dz_elems[2]}, u->img_dgr[2], n->d_h1[0], n->keep_grads ? n->d_dz1 : nullptr,
          PlanePair{u->dz16[3], u->dz_elems[3]}, rows};
      return umma2::launch_umma2("conv2_dgrad", p, rows * P::HC * P::HC, kC1, 4, st);
    }
    default: {
      UmmaState* u = ust(n);
      (void)src; (void)idx; (void)shift;   // the frames were already gathered by conv1_fwd (im2col image)
      WConv1Wgrad p{u->im2col1, PlanePair{u->dz16[3], u->dz_elems[3]}, n->d_part + lt.part_off[0], rows,
                    umma_wgrad_kb(0, rows), g_conv1_tma ? conv1tma::kTileRows : 128};
      return umma_mn::launch_umma_mn("conv1_wgrad", p, kK1, kC1, lt.splits[0], st);
    }
  }
}

// fc1 wgrad + RMSProp + image refresh in one kernel (single-GPU tcgen05 path); must follow fc1_dgrad.
int umma_fc1_wgrad_fused(b200dqn_net* n, int rows, cudaStream_t st, bool keep_grads) {
  UmmaState* u = ust(n);
  const LayerTable& lt = n->lt;
  WFc1WgradFused p{PlanePair{u->h16[2][0], u->h_elems[2]}, PlanePair{u->dz16[0], u->dz_elems[0]},
                   n->d_w + lt.off[3], n->d_s + lt.off[3], keep_grads ? n->d_part + lt.part_off[3] : nullptr,
                   u->img_dgr[0], rows, make_opt_args(n, rows)};
  return umma_mn::launch_umma_mn("fc1_wgrad+opt", p, kFlat, kHidden, 1, st);
}

int umma_fc1_splits(int rows) { return fc1_splits_for(rows); }
bool umma_has_backward() { return true; }
// ---- gather schedule hooks (data-parallel learners, comm_p2p.cuh) ---------------------------------------
int umma_push_h3(b200dqn_net* n, cudaStream_t st) {
  UmmaState* u = ust(n);
  cudaStream_t sN = n->side[3];
  B2_CHECK_CUDA(cudaEventRecord(n->ev[13], st));          // conv3_fwd done: the online net's H3 planes are final
  B2_CHECK_CUDA(cudaStreamWaitEvent(sN, n->ev[13], 0));
  const int rc = comm_push_planes(n, 0, u->h16[2][0], u->h_elems[2], sN);
  if (rc) return rc;
  B2_CHECK_CUDA(cudaEventRecord(n->ev[14], sN));
  return B200DQN_OK;
}
int umma_push_dz4(b200dqn_net* n, cudaStream_t st) {
  UmmaState* u = ust(n);
  return comm_push_planes(n, 1, u->dz16[0], u->dz_elems[0], st);
}
int umma_gather_dz4_ll(b200dqn_net* n, cudaStream_t st) {
  UmmaState* u = ust(n);
  return comm_gather_dz4_ll(n, u->dz16[0], u->dz_elems[0], st, true);   // also waits for the peers' H3 rows
}
int umma_fc1_wgrad_gathered(b200dqn_net* n, cudaStream_t st) {
  WFc1WgradGather p{reinterpret_cast<const __half*>(n->d_xbuf + n->x_h3_off),
                    reinterpret_cast<const __half*>(n->d_xbuf + n->x_dz_off),
                    n->x_h3_parity / 2, n->x_h3_lo, n->x_dz_parity / 2, n->x_dz_lo,
                    n->d_xpush_epoch, n->d_part + n->lt.part_off[3], n->nb * n->world};
  return umma_mn::launch_umma_mn("fc1_wgrad", p, kFlat, kHidden, 1, st);
}

int umma_forward_launches() { return 4; }
int umma_backward_launches() { return 7; }

}  // namespace b200

extern "C" int b200dqn_debug_trace(unsigned long long* host_out, int n) {
  cudaDeviceSynchronize();
  return b200::umma2::read_trace(host_out, n);
}
