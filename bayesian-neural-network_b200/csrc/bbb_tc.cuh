// sm_100a tensor-core plumbing used by the tcgen05 kernels: mbarrier, TMEM allocation, UMMA
// shared-memory / instruction descriptors for kind::tf32, tcgen05.mma / commit / ld, proxy fences,
// and the canonical K-major SWIZZLE_128B tile addressing the CTA's threads write operands with.
//
// Layout facts (checked against cute/arch/mma_sm100_desc.hpp and cute/atom/mma_traits_sm100.hpp of the
// CUTLASS tree vendored in this image; the code below is hand-written PTX, nothing is included from there):
//  * operand tile = [rows][32 fp32] (one 128-byte swizzle row per matrix row), rows grouped by 8 into
//    1024-byte atoms; inside an atom the 16-byte chunk index is XORed with (row & 7)  (Swizzle<3,4,3>).
//  * smem descriptor: bits[0,14) addr>>4, [16,30) LBO>>4 (unused for swizzled K-major, 1), [32,46) SBO>>4
//    (= 1024>>4 between 8-row groups), [46,48) version = 1 (Blackwell), [61,64) layout = 2 (SWIZZLE_128B).
//  * advancing K by 8 tf32 (one MMA) inside the 128-byte row = +32 bytes on the start address.
//  * instruction descriptor (kind::tf32): c_format F32 = 1 @bit4, a/b_format TF32 = 2 @bits7/10, K-major both,
//    N>>3 @bit17, M>>4 @bit24.
//  * accumulator D[M=128][N] in TMEM: row m -> lane m, column n -> column base + n (32-bit each).
#pragma once
#include <stdint.h>

namespace bbb {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// same, for waits that are expected to be long (a warp whose only job is to wait for producers): the suspend-time
// hint lets the hardware park the thread instead of re-issuing the poll, which would steal issue slots
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
  }
}

// ---- fences -----------------------------------------------------------------------------
// generic-proxy smem writes (st.shared by the staging threads) -> visible to the async proxy (tcgen05.mma)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM -------------------------------------------------------------------------------
// one full warp allocates `cols` (power of two >= 32) columns; the base address lands in *smem_dst
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__host__ __device__ constexpr uint32_t tmem_cols_pow2(uint32_t c) {
  return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512;
}

// 8 consecutive accumulator columns of this thread's TMEM lane (warp w reads lanes 32*(w%4)..+31)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float v[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- descriptors ------------------------------------------------------------------------
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);  // start address
  d |= (uint64_t)1u << 16;                  // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024u >> 4) << 32;        // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1u << 46;                  // descriptor version (Blackwell)
  d |= (uint64_t)2u << 61;                  // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// same with operand major-ness: bit 15 = A is MN-major, bit 16 = B is MN-major (0 = K-major)
__host__ __device__ constexpr uint32_t idesc_tf32_major(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
  return idesc_tf32(M, N) | (a_mn << 15) | (b_mn << 16);
}
// MN-major kind::tf32 operands exist in ONE shared-memory format, SWIZZLE_128B_BASE32B (layout type 1; probed on
// B200 with tools/mma_probe.cu, and what CUTLASS' builder states: "for mn-major tf32 operands, SW128_32B is the only
// available smem layout").  The matrix is stored as stacked regions of [K rows][128 bytes = 32 MN elements]; inside
// a row the 32-BYTE chunk index is XORed with (row & 3); swizzle atoms are 4 K rows tall.
//   LBO = byte distance between consecutive 32-element MN groups (regions), SBO = 512 = distance between consecutive
//   4-row K atoms.  Advancing K by 8 (one MMA) = +1024 bytes on the start address.
__device__ __forceinline__ uint64_t smem_desc_mn32(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(512u >> 4) << 32;
  d |= (uint64_t)1u << 46;
  d |= (uint64_t)1u << 61;
  return d;
}
// byte offset of the 16-byte chunk `c16` (0..7) of row `row` inside one such region
__device__ __forceinline__ uint32_t mn32_off(int row, int c16) {
  return (uint32_t)((row << 7) + ((((c16 >> 1) ^ row) & 3) << 5) + ((c16 & 1) << 4));
}

// D[tmem] (+)= A[smem] * B[smem]^T, one K = 8 slice; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- operand tile addressing -------------------------------------------------------------
// byte offset of the 16-byte chunk `chunk` (0..7) of row `row` inside a [rows][32 fp32] SW128 K-major tile
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) {
  return (uint32_t)(((row >> 3) << 10) + ((row & 7) << 7) + ((chunk ^ (row & 7)) << 4));
}
// fp32 -> tf32 with round-to-nearest (ties away, as cvt.rna.tf32.f32): the tensor core itself truncates the low 13
// mantissa bits of whatever it reads, so adding half a TF32 ulp to the bit pattern is the whole conversion -- one
// integer add where cvt.rna compiles to four instructions
__device__ __forceinline__ float to_tf32(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }
__device__ __forceinline__ void st_tile4(uint8_t *tile, int row, int chunk, float a, float b, float c, float d) {
  *reinterpret_cast<float4 *>(tile + sw128_off(row, chunk)) = make_float4(to_tf32(a), to_tf32(b), to_tf32(c), to_tf32(d));
}

}  // namespace tc
}  // namespace bbb
