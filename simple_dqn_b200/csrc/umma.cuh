// umma.cuh — tcgen05 (5th-gen tensor core) building blocks for sm_100a, and the generic
// software-staged implicit-GEMM kernel built on them.
//
//   D[128 x BN] (fp32, TMEM) = A[128 x K] * B[BN x K]^T        A, B: fp16, K-major, 128B-swizzled smem
//
// Operands are produced element-wise by a problem functor (im2col / transposes / u8 frames are
// all "just a functor"), split into fp16 hi + scaled fp16 lo parts and written by all threads
// straight into the canonical UMMA shared-memory layout; one elected thread issues
// tcgen05.mma.  Three MMAs per k-step reproduce fp32 products to ~2^-22:
//     acc0 += A_hi * B_hi            acc1 += A_lo' * B_hi + A_hi * B_lo'      (x' = (x - hi) * 2^11)
//     D = acc0 + acc1 * 2^-11
// (SURVEY §7 "Precision vs the 1e-3 bar": plain fp16 fails the parity bar, the 3-term split is
// indistinguishable from fp32.)  Tensor FLOPs are ~3x the algorithmic FLOPs and still far from
// being the bottleneck at these sizes.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace b200 {
namespace umma {

constexpr int kBM = 128;        // UMMA M (cta_group::1)
constexpr int kBK = 64;         // fp16 elements per k-block = one 128-byte swizzle row
constexpr int kThreads = 256;   // 8 warps: all stage operands; lane quarters x column halves in the epilogue
constexpr float kLoScale = 2048.0f, kLoInv = 1.0f / 2048.0f;

// ---- shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), K-major SW128:
//   [0,14) start>>4 | [16,30) LBO>>4 (=1) | [32,46) SBO>>4 (8 rows * 128 B = 1024 -> 64) |
//   [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  return uint64_t((smem_addr >> 4) & 0x3FFF) | (uint64_t(1) << 16) | (uint64_t(64) << 32) | (uint64_t(1) << 46) |
         (uint64_t(2) << 61);
}

// ---- instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16, fp16 x fp16 -> fp32, both K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4)                      // c_format = F32
         | (0u << 7) | (0u << 10)       // a_format = b_format = F16
         | (0u << 15) | (0u << 16)      // a_major = b_major = K
         | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrives on the mbarrier when every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 8 consecutive fp32 columns: thread t of the warp gets lane (taddr.lane + t)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float v[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// fp32 x[8] -> 16-byte chunks of fp16 hi and scaled fp16 lo
__device__ __forceinline__ void split8(const float x[8], uint4& hi, uint4& lo) {
  __half2 h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
    const float2 back = __half22float2(hh);
    h[i] = hh;
    l[i] = __floats2half2_rn((x[2 * i] - back.x) * kLoScale, (x[2 * i + 1] - back.y) * kLoScale);
  }
  hi = *reinterpret_cast<uint4*>(h);
  lo = *reinterpret_cast<uint4*>(l);
}

// byte offset of 16-byte chunk c (0..7) of row r inside a [rows x 64] fp16 SW128 tile
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return uint32_t((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

// ------------------------------------------------------------------------------------------
// Generic kernel.  grid = (ceil(M/128), ceil(N/BN), Z), block = 256, dynamic smem = smem_bytes<P>().
// Problem P:
//   static constexpr int  kBN;          N tile (multiple of 16, <= 256)
//   static constexpr bool kAExact;      A is exactly representable in fp16 (u8 pixels): no A_lo
//   static constexpr bool kARowMajorThreads / kBRowMajorThreads;
//        true : consecutive threads take consecutive 8-element chunks of one row  (K-contiguous source)
//        false: consecutive threads take consecutive rows of one chunk            (M/N-contiguous source)
//   int M(z), N(z); void krange(z, kb, ke)   (k-block range, in units of 64)
//   void a8(z, m, k0, float v[8]);  void b8(z, n, k0, float v[8]);   rows/cols out of range -> zeros
//   void store8(z, m, n0, const float v[8]);   called only for m < M(z), n0 < N(z) (n0 % 8 == 0)
// ------------------------------------------------------------------------------------------
template <class P>
struct Cfg {
  static constexpr int BN = P::kBN;
  static constexpr int kStages = 3;
  static constexpr uint32_t kABytes = kBM * 128;           // one fp16 [128 x 64] tile
  static constexpr uint32_t kBBytes = BN * 128;
  static constexpr uint32_t kStageBytes = (P::kAExact ? 1 : 2) * kABytes + 2 * kBBytes;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 64 /*barriers*/;
  static constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                        : (2 * BN <= 256) ? 256 : 512;
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128 must be a multiple of 16 in [16,256]");
  static_assert(BN % 16 == 0, "epilogue splits columns in two halves of 8-column chunks");
};

template <class P>
__global__ void __launch_bounds__(kThreads, 1) k_umma_gemm(const P p) {
  using C = Cfg<P>;
  constexpr int BN = C::BN;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_empty[C::kStages];
  __shared__ __align__(8) uint64_t s_done;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int z = blockIdx.z;
  const int M = p.M(z), N = p.N(z);
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * BN;
  if (m0 >= M || n0 >= N) return;
  int kb0, kb1;
  p.krange(z, kb0, kb1);

  // 1024-byte aligned stage buffers (SWIZZLE_128B atoms are 1024 B)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  if (warp == 0) tmem_alloc(&s_tmem, C::kTmemCols);
  if (tid == 32) {
#pragma unroll
    for (int s = 0; s < C::kStages; ++s) mbar_init(&s_empty[s], 1);
    mbar_init(&s_done, 1);
    mbar_fence_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = s_tmem;
  constexpr uint32_t idesc = make_idesc_f16(kBM, BN);

  constexpr int kAChunks = kBM * 8 / kThreads;                 // 16-byte chunks per thread per k-block
  constexpr int kBChunks = (BN * 8 + kThreads - 1) / kThreads;

  for (int kb = kb0; kb < kb1; ++kb) {
    const int it = kb - kb0, s = it % C::kStages;
    if (it >= C::kStages) mbar_wait(&s_empty[s], ((it / C::kStages) - 1) & 1);  // MMAs reading stage s are done
    uint8_t* st_gen = smem_gen + s * C::kStageBytes;
    uint8_t* a_hi = st_gen;
    uint8_t* a_lo = st_gen + C::kABytes;                       // unused when kAExact
    uint8_t* b_hi = st_gen + (P::kAExact ? 1 : 2) * C::kABytes;
    uint8_t* b_lo = b_hi + C::kBBytes;
    const int k0 = kb * kBK;

    float av[kAChunks][8];
#pragma unroll
    for (int i = 0; i < kAChunks; ++i) {
      const int id = tid + i * kThreads;
      const int r = P::kARowMajorThreads ? (id >> 3) : (id % kBM);
      const int c = P::kARowMajorThreads ? (id & 7) : (id / kBM);
      p.a8(z, m0 + r, k0 + c * 8, av[i]);
    }
    float bv[kBChunks][8];
#pragma unroll
    for (int i = 0; i < kBChunks; ++i) {
      const int id = tid + i * kThreads;
      const int r = P::kBRowMajorThreads ? (id >> 3) : (id % BN);
      const int c = P::kBRowMajorThreads ? (id & 7) : (id / BN);
      if (id < BN * 8) p.b8(z, n0 + r, k0 + c * 8, bv[i]);
    }
#pragma unroll
    for (int i = 0; i < kAChunks; ++i) {
      const int id = tid + i * kThreads;
      const int r = P::kARowMajorThreads ? (id >> 3) : (id % kBM);
      const int c = P::kARowMajorThreads ? (id & 7) : (id / kBM);
      uint4 hi, lo;
      split8(av[i], hi, lo);
      *reinterpret_cast<uint4*>(a_hi + sw128_off(r, c)) = hi;
      if (!P::kAExact) *reinterpret_cast<uint4*>(a_lo + sw128_off(r, c)) = lo;
    }
#pragma unroll
    for (int i = 0; i < kBChunks; ++i) {
      const int id = tid + i * kThreads;
      const int r = P::kBRowMajorThreads ? (id >> 3) : (id % BN);
      const int c = P::kBRowMajorThreads ? (id & 7) : (id / BN);
      if (id < BN * 8) {
        uint4 hi, lo;
        split8(bv[i], hi, lo);
        *reinterpret_cast<uint4*>(b_hi + sw128_off(r, c)) = hi;
        *reinterpret_cast<uint4*>(b_lo + sw128_off(r, c)) = lo;
      }
    }
    fence_proxy_async_smem();  // generic-proxy st.shared -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      const uint32_t sa = smem_base + s * C::kStageBytes;
      const uint64_t da_hi = make_desc_sw128(sa);
      const uint64_t da_lo = make_desc_sw128(sa + C::kABytes);
      const uint64_t db_hi = make_desc_sw128(sa + (P::kAExact ? 1 : 2) * C::kABytes);
      const uint64_t db_lo = make_desc_sw128(sa + (P::kAExact ? 1 : 2) * C::kABytes + C::kBBytes);
#pragma unroll
      for (int k = 0; k < kBK / 16; ++k) {                     // UMMA_K = 16 fp16 = 32 bytes = +2 in the address field
        const uint32_t first = (it > 0 || k > 0) ? 1u : 0u;
        mma_f16(tmem, da_hi + 2 * k, db_hi + 2 * k, idesc, first);            // acc0 += A_hi * B_hi
        if (!P::kAExact) {
          mma_f16(tmem + BN, da_lo + 2 * k, db_hi + 2 * k, idesc, first);     // acc1 += A_lo * B_hi
          mma_f16(tmem + BN, da_hi + 2 * k, db_lo + 2 * k, idesc, 1u);        // acc1 += A_hi * B_lo
        } else {
          mma_f16(tmem + BN, da_hi + 2 * k, db_lo + 2 * k, idesc, first);     // acc1 += A_hi * B_lo
        }
      }
      mma_commit(&s_empty[s]);
      if (kb == kb1 - 1) mma_commit(&s_done);
    }
  }

  // ---- epilogue: TMEM -> registers -> functor.  warp w owns lanes [32*(w&3), +32) and column half (w>>2).
  mbar_wait(&s_done, 0);
  fence_after_sync();
  {
    const int q = warp & 3, half = warp >> 2;
    const int m = m0 + q * 32 + lane;
    const uint32_t lane_addr = tmem + (uint32_t(q * 32) << 16);
    constexpr int kColsPerHalf = BN / 2;
#pragma unroll
    for (int c = 0; c < kColsPerHalf; c += 8) {
      const int col = half * kColsPerHalf + c;
      float a0[8], a1[8];
      tmem_ld8(lane_addr + col, a0);
      tmem_ld8(lane_addr + BN + col, a1);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) a0[j] = fmaf(a1[j], kLoInv, a0[j]);
      if (m < M && n0 + col < N) p.store8(z, m, n0 + col, a0);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    fence_after_sync();
    tmem_dealloc(tmem, C::kTmemCols);
  }
}

template <class P>
static int launch_umma(const char* label, const P& p, int M, int N, int Z, cudaStream_t st) {
  using C = Cfg<P>;
  static bool configured = false;
  if (!configured) {
    B2_CHECK_CUDA(cudaFuncSetAttribute(k_umma_gemm<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    configured = true;
  }
  dim3 grid((M + kBM - 1) / kBM, (N + C::BN - 1) / C::BN, Z);
  k_umma_gemm<P><<<grid, kThreads, C::kSmemBytes, st>>>(p);
  B2_LAUNCH_CHECK();
  B2_PROF(label, st);
  return B200DQN_OK;
}

}  // namespace umma
}  // namespace b200
