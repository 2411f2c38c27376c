// umma.cuh — tcgen05 (5th-gen tensor core) building blocks for sm_100a shared by the kernels of
// umma2.cuh (K-major operands: forward, dgrad) and umma_mn.cuh (MN-major operands: wgrad).
//
//   D[128 x BN] (fp32, TMEM) = A[128 x K] * B[BN x K]^T        A, B: fp16 in SWIZZLE_128B shared memory
//
// Every fp32 operand x is split ONCE, by its producer, into hi = fp16(x) and lo' = fp16((x - hi) * 2^11);
// a k-step issues  [acc0 | acc1] += A_hi x [B_hi ; B_lo']  (one MMA with N = 2*BN: the hi and lo tiles
// are adjacent in shared memory) and  acc1 += A_lo' x B_hi;  the epilogue returns acc0 + acc1 * 2^-11.
// That reproduces fp32 products to ~2^-22 (SURVEY §7 "Precision vs the 1e-3 bar": plain fp16 fails the
// parity bar, the split is indistinguishable from fp32) for 1.5-2x the tensor work of plain fp16.
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

// MN-major SWIZZLE_128B descriptor: LBO = byte stride between 64-element MN chunks, SBO = byte stride
// between groups of 8 k rows (1024 when rows are packed).  A [64 k-rows x 128 B] sub-tile written in the K-major
// SW128 pattern with "row = k" IS this layout — which is why the row-oriented fc1 weight image (rows = flat index,
// 64 hidden units per 128-byte row) serves the dgrad as a K-major operand and the forward as an MN-major one.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  return uint64_t((smem_addr >> 4) & 0x3FFF) | (uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16) | (uint64_t(64) << 32) |
         (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
constexpr uint32_t kIdescAMn = 1u << 15, kIdescBMn = 1u << 16;   // a_major / b_major = MN

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

}  // namespace umma
}  // namespace b200
