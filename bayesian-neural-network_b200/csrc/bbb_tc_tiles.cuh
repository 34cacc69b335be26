// Tile-level building blocks shared by the tcgen05 weight-sampling kernels (bbb_linear_tc.cu: batch-tiled
// kernels for B > 128; bbb_linear_sk.cu: stream-K / balanced kernels for B <= 128): weight quads and their
// sampling, activation row loads, UMMA block issue, and the TMEM -> swizzled smem -> coalesced global drain.
#pragma once
#include "bbb_common.cuh"
#include "bbb_kernels.h"
#include "bbb_tc.cuh"

namespace bbb {
namespace tcx {

using namespace tc;

constexpr int NT = 256;          // threads per CTA
constexpr int BM = 128;          // UMMA M
constexpr int BK = 32;           // fp32 elements per 128-byte swizzle row
constexpr int A_TILE = BM * 128; // bytes of one [128][32] operand tile

template <int BN, int SG>
struct Smem {
  static constexpr int kB = BN * 128;                       // bytes of one [BN][32] tile
  static constexpr int kStage = SG * A_TILE + SG * kB;      // one pipeline stage
  static constexpr int kTiles = 2 * kStage;
  static constexpr int kDyn = kTiles + 1024;                // + alignment slack
  static constexpr uint32_t kTmemCols = tmem_cols_pow2(SG * BN);
};

struct Ctl {  // small static shared state
  uint64_t bar[3];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint8_t *align1024(uint8_t *p) {
  return reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}

__device__ __forceinline__ void ctl_setup(Ctl &c, uint32_t tmem_cols) {
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&c.tmem_base), tmem_cols);
  if (threadIdx.x == 32) {
    mbar_init(smem_u32(&c.bar[0]), 1);
    mbar_init(smem_u32(&c.bar[1]), 1);
    mbar_init(smem_u32(&c.bar[2]), 1);
    mbar_fence_init();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
}
__device__ __forceinline__ void ctl_teardown(Ctl &c, uint32_t tmem_cols) {
  tc_fence_before_sync();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) tmem_dealloc(c.tmem_base, tmem_cols);
}

// ---- one quad (4 consecutive k/i elements of one weight row) ------------------------------------
struct Quad {
  float mu[4], sg[4], rho[4];
};
__device__ __forceinline__ void load_quad(const LinArgs &a, int64_t e, bool need_sigma, Quad &q) {
  const float4 m = __ldg(reinterpret_cast<const float4 *>(a.w_mu + e));
  q.mu[0] = m.x; q.mu[1] = m.y; q.mu[2] = m.z; q.mu[3] = m.w;
  if (need_sigma) {
    const float4 r = __ldg(reinterpret_cast<const float4 *>(a.w_rho + e));
    q.rho[0] = r.x; q.rho[1] = r.y; q.rho[2] = r.z; q.rho[3] = r.w;
#pragma unroll
    for (int j = 0; j < 4; ++j) q.sg[j] = softplus_fast(q.rho[j]);
  }
}
// sigmoid(rho) for the rho-gradient epilogue (SFU exp + fast divide)
__device__ __forceinline__ float sigmoid_fast(float rho) {
  const float t = __expf(-fabsf(rho));
  const float r = __fdividef(1.0f, 1.0f + t);
  return rho >= 0.0f ? r : t * r;
}
// eps and w of sample s for the quad at linear element e (e % 4 == 0)
__device__ __forceinline__ void sample_quad(const LinArgs &a, int s, int64_t e, const Quad &q, bool sample,
                                            float ep[4], float w[4]) {
  if (sample) {
    if (a.eps_w) {
      const float4 t = __ldg(reinterpret_cast<const float4 *>(a.eps_w + (int64_t)s * a.out * a.in + e));
      ep[0] = t.x; ep[1] = t.y; ep[2] = t.z; ep[3] = t.w;
    } else {
      philox_normal4(a.rng, a.rng.tensor_w, a.rng.sample_base + (uint32_t)s, (uint32_t)(e >> 2), ep);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = __fadd_rn(q.mu[j], __fmul_rn(q.sg[j], ep[j]));
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) { ep[j] = 0.0f; w[j] = q.mu[j]; }
  }
}
__device__ __forceinline__ void bias_elem(const LinArgs &a, int s, int64_t o, bool sample, bool need_sigma, float &b,
                                          float &sg, float &ep) {
  const float mu = __ldg(a.b_mu + o);
  sg = (sample || need_sigma) ? softplus_f(__ldg(a.b_rho + o)) : 0.0f;
  ep = 0.0f;
  if (sample)
    ep = a.eps_b ? __ldg(a.eps_b + (int64_t)s * a.out + o)
                 : philox_normal1(a.rng, a.rng.tensor_b, a.rng.sample_base + (uint32_t)s, (uint64_t)o);
  b = sample ? __fadd_rn(mu, __fmul_rn(sg, ep)) : mu;
}

// 4 consecutive floats of a row-major [rows][ld] matrix at (r, c), zero outside; vector load when allowed
__device__ __forceinline__ float4 ld_row4(const float *__restrict__ p, int64_t r, int64_t c, int64_t rows, int64_t ld,
                                          bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r < rows && c < ld) {
    if (vec && c + 3 < ld) {
      v = __ldg(reinterpret_cast<const float4 *>(p + r * ld + c));
    } else {
      const float *q = p + r * ld + c;
      v.x = __ldg(q);
      if (c + 1 < ld) v.y = __ldg(q + 1);
      if (c + 2 < ld) v.z = __ldg(q + 2);
      if (c + 3 < ld) v.w = __ldg(q + 3);
    }
  }
  return v;
}
// same for 16-byte aligned rows (ld % 4 == 0, c % 4 == 0, aligned base), branch-free: the address is clamped into the
// matrix, the load is unconditional and the result is zeroed afterwards -- so a batch of these issues back to back
__device__ __forceinline__ float4 ld_row4_al(const float *__restrict__ p, int64_t r, int64_t c, int64_t rows, int64_t ld) {
  const bool ok = r < rows && c < ld;
  const float4 v = __ldg(reinterpret_cast<const float4 *>(p + (ok ? r * ld + c : 0)));
  return ok ? v : make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ float4 relu4(float4 v) {
  return make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
}
__device__ __forceinline__ float4 mask4(float4 v, float4 m) {
  return make_float4(m.x > 0.f ? v.x : 0.f, m.y > 0.f ? v.y : 0.f, m.z > 0.f ? v.z : 0.f, m.w > 0.f ? v.w : 0.f);
}

// issue the 4 K=8 MMAs of one 32-wide k block for one accumulator
__device__ __forceinline__ void issue_block(uint32_t tmem_d, uint32_t a_saddr, uint32_t b_saddr, uint32_t idesc,
                                            bool first) {
  const uint64_t da = smem_desc_sw128(a_saddr), db = smem_desc_sw128(b_saddr);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) mma_tf32(tmem_d, da + 2u * kk, db + 2u * kk, idesc, (first && kk == 0) ? 0u : 1u);
}


__device__ __forceinline__ void red_add_v4(float *addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// Drain the SG accumulators [128][BN] of a CTA from TMEM and write them row-major to global memory with
// coalesced 16-byte accesses.  TMEM hands every thread one ROW (lane) at a time, global memory wants a warp
// on consecutive COLUMNS, so the tile goes through shared memory (the operand buffers, free once the MMAs
// have completed) in a layout whose 16-byte chunk index is XORed with (row & 7): both the row-wise writes
// and the column-wise reads are bank-conflict free.  Split-K partial sums are combined with
// red.global.add.v4.f32 (one request per 16 bytes instead of one per float).
// kBar == 0: all NT threads of the CTA take part (__syncthreads); kBar > 0: the first 256 threads do, meeting
// on named barrier kBar (warp-specialised kernels whose extra warps never enter).
template <int kBar>
__device__ __forceinline__ void drain_sync() {
  if constexpr (kBar == 0) __syncthreads();
  else asm volatile("bar.sync %0, 256;" ::"n"(kBar) : "memory");
}
template <int BN, int SG, int kBar = 0, class Bias>
__device__ __forceinline__ void drain_tile(uint8_t *tiles, uint32_t tmem, int ns, bool have_acc, float *dst,
                                           int64_t sample_stride, int64_t row0, int64_t rows, int64_t col0,
                                           int64_t ld, bool vec_ok, bool atomic, float scale, Bias bias,
                                           const float *mask_src = nullptr) {
  constexpr int NCH = BN / 4, CH = ((NCH + 7) / 8) * 8, PITCH = CH * 16, HALF = BN / 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q4 = warp & 3, half = warp >> 2, row = q4 * 32 + lane;
#pragma unroll
  for (int s = 0; s < SG; ++s) {
    if (s >= ns) break;
    uint8_t *Ys = tiles + s * BM * PITCH + row * PITCH;
#pragma unroll
    for (int c0 = 0; c0 < HALF; c0 += 8) {
      const int col = half * HALF + c0;
      float v[8];
      if (have_acc) {
        tmem_ld8(tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(s * BN + col), v);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.0f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaf(scale, v[j], bias(s, col + j));
      const int c = col >> 2;
      *reinterpret_cast<float4 *>(Ys + (((c) ^ (row & 7)) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4 *>(Ys + (((c + 1) ^ (row & 7)) << 4)) = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
  tc_fence_before_sync();
  drain_sync<kBar>();
#pragma unroll
  for (int s = 0; s < SG; ++s) {
    if (s >= ns) break;
    float *out = dst + (int64_t)s * sample_stride;
    for (int idx = tid; idx < BM * NCH; idx += NT) {
      const int r = idx / NCH, c = idx % NCH;
      const int64_t gr = row0 + r, gc = col0 + c * 4;
      if (gr >= rows || gc >= ld) continue;
      float4 v = *reinterpret_cast<const float4 *>(tiles + s * BM * PITCH + r * PITCH + ((c ^ (r & 7)) << 4));
      float *p = out + gr * ld + gc;
      if (mask_src) {  // multiply by (mask_src > 0) at the same [sample][row][column] position (linear: fine for partials)
        const float *m = mask_src + (int64_t)s * sample_stride + gr * ld + gc;
        v.x = m[0] > 0.f ? v.x : 0.f;
        if (gc + 1 < ld) v.y = m[1] > 0.f ? v.y : 0.f;
        if (gc + 2 < ld) v.z = m[2] > 0.f ? v.z : 0.f;
        if (gc + 3 < ld) v.w = m[3] > 0.f ? v.w : 0.f;
      }
      if (vec_ok && gc + 3 < ld) {
        if (atomic) red_add_v4(p, v);
        else *reinterpret_cast<float4 *>(p) = v;
      } else {
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (gc + j < ld) {
            if (atomic) atomicAdd(p + j, e[j]);
            else p[j] = e[j];
          }
      }
    }
  }
}

}  // namespace tcx
}  // namespace bbb
