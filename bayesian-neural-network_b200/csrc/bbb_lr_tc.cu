// Local-reparameterisation Bayesian linear layer on tcgen05 kind::tf32, batches of at most 128 rows.
//
// BayesianLinearLR.forward (networks.py:116-138) is a PAIRED contraction, gamma = x mu and v = x^2 sigma^2, followed
// by out = gamma + sqrt(v) eps_a + (mu_b + sigma_b eps_b).  Weights are [in, out] as in the reference, so a weight
// tile [32 k][32 o] with o contiguous is an MN-major B operand as it stands (SWIZZLE_128B_BASE32B, bbb_tc.cuh): mu
// is a raw copy, sigma^2 = softplus(rho)^2 is formed on the way.  Nothing is sampled per weight, so the kernel is
// bound by staging, not by ALU work.
//
//   lr_fwd_sk_kernel     stream-K over units [32 o x 32 k] (same work split and warp roles as bbb_linear_sk.cu): the
//                        8 producer warps stage x, x^2 (K-major) and mu, sigma^2 (MN-major); the MMA warp runs the two
//                        contractions into two TMEM accumulators; partial tiles are added into y (gamma) and delta (v)
//                        with red.add.  The closed-form KL (networks.py:109-114) is summed on the way, once per call.
//   lr_epilogue_kernel   elementwise over [S, B, out]: delta = sqrt(v), y = gamma + delta eps_a + bias sample, with
//                        eps from Philox (same coordinates as the fp32 path: bbb_lr_fma.cu) or from memory.
#include "bbb_tc_tiles.cuh"

namespace bbb {
namespace {

using namespace tc;
using namespace tcx;

constexpr int NPW = 8, NTP = 256, NT2 = NTP + 32, NS = 2, kCtaPerSm = 2;
constexpr int BN = 32;                       // output columns per unit (one MN-major region)
constexpr int B_TILE = 32 * 128;             // bytes of one [32 k][32 o] weight tile
constexpr int kStage = 2 * A_TILE + 2 * B_TILE;   // x | x^2 | mu | sigma^2
constexpr int kTiles = NS * kStage;
constexpr int kDyn = kTiles + 1024;
constexpr uint32_t kTmemCols = 64;           // gamma: [0, 32), v: [32, 64)

struct Ctl2 {
  uint64_t full[NS], empty[NS], acc;
  uint32_t tmem_base;
};
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_producers() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// KL(N(mu, sigma) || N(0, sigma_p)) of one element; c0 = 2 log sigma_p - 1
__device__ __forceinline__ float kl_elem_fast(float mu, float sg, float c0, float inv_sp2) {
  return 0.5f * (c0 - 2.0f * 0.6931471805599453f * lg2_approx(sg) + (sg * sg + mu * mu) * inv_sp2);
}

template <bool kSample>
__global__ void __launch_bounds__(NT2, kCtaPerSm) lr_fwd_sk_kernel(const LrArgs a_in, int nkb, int o_tiles, int total) {
  extern __shared__ uint8_t dsm[];
  __shared__ Ctl2 ctl;
  __shared__ float red[64];
  __shared__ float bias_s[BN];
  LrArgs a = a_in;
  uint8_t *tiles = align1024(dsm);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool relu = a.flags & BBB_F_RELU_IN, calc_kl = a.flags & BBB_F_LOGPROB;
  const int u0 = (int)((int64_t)blockIdx.x * total / gridDim.x), u1 = (int)((int64_t)(blockIdx.x + 1) * total / gridDim.x);

  if (warp == NPW) tmem_alloc(smem_u32(&ctl.tmem_base), kTmemCols);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      mbar_init(smem_u32(&ctl.full[s]), NPW);
      mbar_init(smem_u32(&ctl.empty[s]), 1);
    }
    mbar_init(smem_u32(&ctl.acc), 1);
    mbar_fence_init();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = ctl.tmem_base;
  pdl_wait();
  rng_resolve(a.rng);

  constexpr uint32_t idesc = idesc_tf32_major(BM, BN, 0, 1);   // A = x (K-major), B = weights (MN-major)
  float kl = 0.0f;

  if (warp == NPW) {
    if (lane == 0) {
      int it = 0;
      for (int u = u0; u < u1;) {
        const int tile = u / nkb, kb0 = u - tile * nkb, kb1 = min(nkb, kb0 + (u1 - u));
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int stage = it % NS;
          mbar_wait_parked(smem_u32(&ctl.full[stage]), (uint32_t)((it / NS) & 1));
          tc_fence_after_sync();
          const uint32_t As = smem_u32(tiles + stage * kStage), Bs = As + 2 * A_TILE;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t acc = (kb == kb0 && kk == 0) ? 0u : 1u;
            mma_tf32(tmem, smem_desc_sw128(As) + 2u * kk, smem_desc_mn32(Bs + kk * 1024, B_TILE), idesc, acc);
            if (kSample)
              mma_tf32(tmem + 32, smem_desc_sw128(As + A_TILE) + 2u * kk, smem_desc_mn32(Bs + B_TILE + kk * 1024, B_TILE),
                       idesc, acc);
          }
          mma_commit(smem_u32(&ctl.empty[stage]));
        }
        mma_commit(smem_u32(&ctl.acc));
        u += kb1 - kb0;
      }
    }
    __syncwarp();
  } else {
    const float c0 = 2.0f * logf(a.sigma_p) - 1.0f, inv_sp2 = 1.0f / (a.sigma_p * a.sigma_p);
    const int wrow = tid >> 3, chunk = tid & 7;
    const uint32_t x_off = sw128_off(wrow, chunk);     // activation rows wrow + 32 j: + 4096 j
    const uint32_t w_off = mn32_off(wrow, chunk);      // weight row k = wrow of the unit, columns 4 chunk .. +3
    int it = 0, seg = 0;
    for (int u = u0; u < u1; ++seg) {
      const int tile = u / nkb, kb0 = u - tile * nkb, kb1 = min(nkb, kb0 + (u1 - u));
      const int s = tile / o_tiles, ot = tile - s * o_tiles;
      const int64_t o0 = (int64_t)ot * BN;
      const bool kl_seg = calc_kl && s == 0;             // every weight is visited once per sample: count it for s = 0
      const int64_t oc = o0 + chunk * 4;
      const bool o_ok = oc < a.out;
      const float *xs = a.x + (int64_t)s * a.x_sstride + chunk * 4;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int stage = it % NS;
        if (it >= NS) mbar_wait(smem_u32(&ctl.empty[stage]), (uint32_t)(((it / NS) - 1) & 1));
        uint8_t *As = tiles + stage * kStage, *Bs = As + 2 * A_TILE;
        const int kbase = kb * BK;
        // loads first
        const bool col_ok = kbase + chunk * 4 < a.in;
        float4 xv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int b = wrow + 32 * j;
          const bool ok = col_ok && b < a.B;
          const float4 v = __ldg(reinterpret_cast<const float4 *>(ok ? xs + (int64_t)b * a.in + kbase : a.x));
          xv[j] = ok ? v : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const int64_t i = kbase + wrow;
        const bool w_ok = o_ok && i < a.in;
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f), r = m;
        const int nv = w_ok ? (int)min((int64_t)4, a.out - oc) : 0;     // valid columns of this quad
        if (nv == 4 && a.vec_out) {
          m = __ldg(reinterpret_cast<const float4 *>(a.w_mu + i * a.out + oc));
          if (kSample || kl_seg) r = __ldg(reinterpret_cast<const float4 *>(a.w_rho + i * a.out + oc));
        } else if (nv > 0) {   // rows of [in, out] that are not 16-byte multiples (e.g. a 10-class head)
          const float *pm = a.w_mu + i * a.out + oc, *pr = a.w_rho + i * a.out + oc;
          m.x = __ldg(pm); r.x = __ldg(pr);
          if (nv > 1) { m.y = __ldg(pm + 1); r.y = __ldg(pr + 1); }
          if (nv > 2) { m.z = __ldg(pm + 2); r.z = __ldg(pr + 2); }
          if (nv > 3) { m.w = __ldg(pm + 3); r.w = __ldg(pr + 3); }
        }
        // weights
        float4 v2 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nv > 0 && (kSample || kl_seg)) {
          const float sg0 = softplus_fast(r.x), sg1 = softplus_fast(r.y), sg2 = softplus_fast(r.z), sg3 = softplus_fast(r.w);
          v2 = make_float4(sg0 * sg0, nv > 1 ? sg1 * sg1 : 0.f, nv > 2 ? sg2 * sg2 : 0.f, nv > 3 ? sg3 * sg3 : 0.f);
          if (kl_seg) {
            kl += kl_elem_fast(m.x, sg0, c0, inv_sp2);
            if (nv > 1) kl += kl_elem_fast(m.y, sg1, c0, inv_sp2);
            if (nv > 2) kl += kl_elem_fast(m.z, sg2, c0, inv_sp2);
            if (nv > 3) kl += kl_elem_fast(m.w, sg3, c0, inv_sp2);
          }
        }
        // operands are rounded to TF32 (cvt.rna) rather than left to the tensor core's truncation: the variance path
        // divides by delta in the backward, which amplifies a truncation bias in x^2 sigma^2
        *reinterpret_cast<float4 *>(Bs + w_off) = make_float4(to_tf32(m.x), to_tf32(m.y), to_tf32(m.z), to_tf32(m.w));
        if (kSample)
          *reinterpret_cast<float4 *>(Bs + B_TILE + w_off) = make_float4(to_tf32(v2.x), to_tf32(v2.y), to_tf32(v2.z), to_tf32(v2.w));
        // activations and their squares
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = relu ? relu4(xv[j]) : xv[j];
          *reinterpret_cast<float4 *>(As + x_off + 4096 * j) = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
          if (kSample)
            *reinterpret_cast<float4 *>(As + A_TILE + x_off + 4096 * j) =
                make_float4(to_tf32(v.x * v.x), to_tf32(v.y * v.y), to_tf32(v.z * v.z), to_tf32(v.w * v.w));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&ctl.full[stage]));
      }
      // the segment that starts a tile at k = 0 owns the tile's bias sample (mu_b + sigma_b eps_b) and, for sample 0,
      // the bias KL
      if (tid < BN) {
        float bv = 0.0f;
        if (kb0 == 0 && o0 + tid < a.out) {
          const float bmu = __ldg(a.b_mu + o0 + tid);
          bv = bmu;
          if (kSample || kl_seg) {
            const float bsg = softplus_fast(__ldg(a.b_rho + o0 + tid));
            if (kSample) {
              const float eb = a.eps_b ? __ldg(a.eps_b + (int64_t)s * a.out + o0 + tid)
                                       : philox_normal1(a.rng, a.rng.tensor_b, a.rng.sample_base + (uint32_t)s, (uint64_t)(o0 + tid));
              bv = fmaf(bsg, eb, bmu);
            }
            if (kl_seg) kl += kl_elem_fast(bmu, bsg, c0, inv_sp2);
          }
        }
        bias_s[tid] = bv;
      }
      mbar_wait_parked(smem_u32(&ctl.acc), (uint32_t)(seg & 1));
      tc_fence_after_sync();
      bar_producers();   // bias_s
      const bool atomic = !(kb0 == 0 && kb1 == nkb);
      drain_tile<BN, 1, 1>(tiles, tmem, 1, true, a.y + (int64_t)s * a.B * a.out, a.B * a.out, 0, a.B, o0, a.out, a.vec_out,
                           atomic, 1.0f, [&](int, int c) { return bias_s[c]; });
      if (kSample) {
        bar_producers();
        drain_tile<BN, 1, 1>(tiles, tmem + 32, 1, true, a.delta + (int64_t)s * a.B * a.out, a.B * a.out, 0, a.B, o0, a.out,
                             a.vec_out, atomic, 1.0f, [](int, int) { return 0.0f; });
      }
      tc_fence_before_sync();
      bar_producers();
      u += kb1 - kb0;
    }
  }
  pdl_launch_dependents();
  __syncthreads();
  if (calc_kl) block_sum2_atomic(kl, 0.0f, red, a.kl, nullptr);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == NPW) tmem_dealloc(tmem, kTmemCols);
}

// out = gamma' + sqrt(v) eps_a in place: on entry y holds gamma' = x mu + bias sample and delta holds v = x^2 sigma^2.
__global__ void __launch_bounds__(256) lr_epilogue_kernel(const LrArgs a_in) {
  pdl_launch_dependents();
  pdl_wait();
  LrArgs a = a_in;
  rng_resolve(a.rng);
  const int per = (int)((a.B * a.out) >> 2);                       // quads per sample
  for (int q = blockIdx.x * 256 + threadIdx.x; q < per * a.S; q += gridDim.x * 256) {
    const int s = q / per;
    const int64_t idx = (int64_t)(q - s * per) << 2, off = (int64_t)s * a.B * a.out + idx;
    float4 g = *reinterpret_cast<float4 *>(a.y + off);
    const float4 v = *reinterpret_cast<const float4 *>(a.delta + off);
    const float4 d = make_float4(sqrtf(fmaxf(v.x, 0.f)), sqrtf(fmaxf(v.y, 0.f)), sqrtf(fmaxf(v.z, 0.f)), sqrtf(fmaxf(v.w, 0.f)));
    float e[4];
    if (a.eps_a) {
      const float4 t = __ldg(reinterpret_cast<const float4 *>(a.eps_a + off));
      e[0] = t.x; e[1] = t.y; e[2] = t.z; e[3] = t.w;
    } else {
      philox_normal4(a.rng, a.rng.tensor_w, a.rng.sample_base + (uint32_t)s, (uint32_t)(idx >> 2), e);
    }
    g.x = fmaf(d.x, e[0], g.x); g.y = fmaf(d.y, e[1], g.y); g.z = fmaf(d.z, e[2], g.z); g.w = fmaf(d.w, e[3], g.w);
    *reinterpret_cast<float4 *>(a.delta + off) = d;
    *reinterpret_cast<float4 *>(a.y + off) = g;
  }
}

// the same, element by element, for outputs whose rows are not 16-byte multiples (Philox per element, as the fp32 path)
__global__ void __launch_bounds__(256) lr_epilogue_scalar_kernel(const LrArgs a_in) {
  pdl_launch_dependents();
  pdl_wait();
  LrArgs a = a_in;
  rng_resolve(a.rng);
  const int per = (int)(a.B * a.out);
  for (int q = blockIdx.x * 256 + threadIdx.x; q < per * a.S; q += gridDim.x * 256) {
    const int s = q / per;
    const int64_t idx = q - s * per, off = (int64_t)s * per + idx;
    const float d = sqrtf(fmaxf(a.delta[off], 0.f));
    const float e = a.eps_a ? __ldg(a.eps_a + off)
                            : philox_normal1(a.rng, a.rng.tensor_w, a.rng.sample_base + (uint32_t)s, (uint64_t)idx);
    a.y[off] = fmaf(d, e, a.y[off]);
    a.delta[off] = d;
  }
}

// ==================================================================================================
// backward.  Three kernels:
//   lr_dv_kernel        elementwise: dV = dz eps_a / (2 delta) (0 where delta == 0), written over delta IN PLACE (the
//                       forward's delta has no other use in the backward); column sums of dz -> bias gradients.
//   lr_dgrad_sk_kernel  dX = dz mu^T + 2 x (dV (sigma^2)^T): the forward's stream-K skeleton with A = dz, dV tiles
//                       and B = mu, sigma^2 tiles, which for this product are K-major as they lie in memory.
//   lr_wgrad_tc_kernel  G1 = x^T dz, G2 = (x^2)^T dV summed over samples AND batch rows (K = S B, chunks of 32 rows),
//                       every operand MN-major in its natural orientation; epilogue
//                       grad_mu = G1 + g mu / sp^2,  grad_rho = sigmoid(rho) (2 sigma G2 + g (sigma / sp^2 - 1 / sigma)).
// ==================================================================================================
__device__ __forceinline__ void eps_a_quad(const LrArgs &a, int s, int64_t idx, float e[4]) {
  if (a.eps_a) {
    const float4 t = __ldg(reinterpret_cast<const float4 *>(a.eps_a + (int64_t)s * a.B * a.out + idx));
    e[0] = t.x; e[1] = t.y; e[2] = t.z; e[3] = t.w;
  } else {
    philox_normal4(a.rng, a.rng.tensor_w, a.rng.sample_base + (uint32_t)s, (uint32_t)(idx >> 2), e);
  }
}

__global__ void __launch_bounds__(1024) lr_dv_kernel(const LrArgs a_in, float *dv_out) {
  __shared__ float part[128][33];
  pdl_launch_dependents();
  pdl_wait();
  LrArgs a = a_in;
  rng_resolve(a.rng);
  const int tid = threadIdx.x, br = tid >> 3, chunk = tid & 7;
  const int64_t o = (int64_t)blockIdx.x * 32 + chunk * 4;
  const bool o_ok = o < a.out, wgrad = !(a.flags & BBB_F_NO_WGRAD);
  float gbmu = 0.0f, gbrho = 0.0f;   // threads tid < 32: column blockIdx.x * 32 + tid
  for (int s = 0; s < a.S; ++s) {
    const int64_t base = (int64_t)s * a.B * a.out;
    float cs[4] = {0.f, 0.f, 0.f, 0.f};
    if (o_ok) {
      for (int64_t b = br; b < a.B; b += 128) {
        const int64_t idx = b * a.out + o;
        const float4 dz = __ldg(reinterpret_cast<const float4 *>(a.dy + base + idx));
        const float4 d = __ldg(reinterpret_cast<const float4 *>(a.delta_in + base + idx));
        float e[4];
        eps_a_quad(a, s, idx, e);
        float4 dv;
        dv.x = d.x > 0.0f ? dz.x * e[0] / (2.0f * d.x) : 0.0f;
        dv.y = d.y > 0.0f ? dz.y * e[1] / (2.0f * d.y) : 0.0f;
        dv.z = d.z > 0.0f ? dz.z * e[2] / (2.0f * d.z) : 0.0f;
        dv.w = d.w > 0.0f ? dz.w * e[3] / (2.0f * d.w) : 0.0f;
        *reinterpret_cast<float4 *>(dv_out + base + idx) = dv;
        cs[0] += dz.x; cs[1] += dz.y; cs[2] += dz.z; cs[3] += dz.w;
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) part[br][chunk * 4 + j] = cs[j];
    __syncthreads();
    if (tid < 32) {
      float colsum = 0.0f;
#pragma unroll 8
      for (int r = 0; r < 128; ++r) colsum += part[r][tid];
      const int64_t oc = (int64_t)blockIdx.x * 32 + tid;
      if (oc < a.out) {
        const float eb = a.eps_b ? __ldg(a.eps_b + (int64_t)s * a.out + oc)
                                 : philox_normal1(a.rng, a.rng.tensor_b, a.rng.sample_base + (uint32_t)s, (uint64_t)oc);
        gbmu += colsum;
        gbrho += colsum * eb;
      }
    }
  }
  const int64_t oc = (int64_t)blockIdx.x * 32 + tid;
  if (wgrad && tid < 32 && oc < a.out) {
    const bool accum = a.flags & BBB_F_ACCUM;
    const float osc = a.out_scale_dev ? __ldg(a.out_scale_dev) : 1.0f;
    const float gk = (a.flags & BBB_F_LOGPROB) ? a.g_kl * (a.g_kl_dev ? __ldg(a.g_kl_dev) : 1.0f) : 0.0f;
    const float inv_sp2 = 1.0f / (a.sigma_p * a.sigma_p);
    const float mu = a.b_mu[oc], sg = softplus_f(a.b_rho[oc]);
    const float gm = fmaf(gk * mu, inv_sp2, gbmu);
    const float gr = -expm1f(-sg) * (gbrho + gk * (sg * inv_sp2 - 1.0f / sg));
    a.g_b_mu[oc] = accum ? fmaf(osc, gm, a.g_b_mu[oc]) : osc * gm;
    a.g_b_rho[oc] = accum ? fmaf(osc, gr, a.g_b_rho[oc]) : osc * gr;
  }
}

// dX_s[b][i] += sum_{o in segment} dz_s[b][o] mu[i][o] + 2 x_s[b][i] sum_o dV_s[b][o] sigma^2[i][o]
__global__ void __launch_bounds__(NT2, kCtaPerSm) lr_dgrad_sk_kernel(const LrArgs a_in, const float *dv, int nkb, int i_tiles,
                                                                      int total) {
  extern __shared__ uint8_t dsm[];
  __shared__ Ctl2 ctl;
  LrArgs a = a_in;
  uint8_t *tiles = align1024(dsm);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool relu = a.flags & BBB_F_RELU_IN, dx_preact = a.flags & BBB_F_DX_PREACT;
  const int u0 = (int)((int64_t)blockIdx.x * total / gridDim.x), u1 = (int)((int64_t)(blockIdx.x + 1) * total / gridDim.x);

  if (warp == NPW) tmem_alloc(smem_u32(&ctl.tmem_base), kTmemCols);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      mbar_init(smem_u32(&ctl.full[s]), NPW);
      mbar_init(smem_u32(&ctl.empty[s]), 1);
    }
    mbar_init(smem_u32(&ctl.acc), 1);
    mbar_fence_init();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = ctl.tmem_base;
  pdl_wait();
  constexpr uint32_t idesc = idesc_tf32(BM, BN);   // both operands K-major
  const float dsc = ((a.flags & BBB_F_SCALE_DX) && a.out_scale_dev) ? __ldg(a.out_scale_dev) : 1.0f;

  if (warp == NPW) {
    if (lane == 0) {
      int it = 0;
      for (int u = u0; u < u1;) {
        const int tile = u / nkb, kb0 = u - tile * nkb, kb1 = min(nkb, kb0 + (u1 - u));
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int stage = it % NS;
          mbar_wait_parked(smem_u32(&ctl.full[stage]), (uint32_t)((it / NS) & 1));
          tc_fence_after_sync();
          const uint32_t As = smem_u32(tiles + stage * kStage), Bs = As + 2 * A_TILE;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t acc = (kb == kb0 && kk == 0) ? 0u : 1u;
            mma_tf32(tmem, smem_desc_sw128(As) + 2u * kk, smem_desc_sw128(Bs) + 2u * kk, idesc, acc);
            mma_tf32(tmem + 32, smem_desc_sw128(As + A_TILE) + 2u * kk, smem_desc_sw128(Bs + B_TILE) + 2u * kk, idesc, acc);
          }
          mma_commit(smem_u32(&ctl.empty[stage]));
        }
        mma_commit(smem_u32(&ctl.acc));
        u += kb1 - kb0;
      }
    }
    __syncwarp();
  } else {
    const int wrow = tid >> 3, chunk = tid & 7;
    const uint32_t t_off = sw128_off(wrow, chunk);   // rows wrow + 32 j of a 128-row tile: + 4096 j
    int it = 0, seg = 0;
    for (int u = u0; u < u1; ++seg) {
      const int tile = u / nkb, kb0 = u - tile * nkb, kb1 = min(nkb, kb0 + (u1 - u));
      const int s = tile / i_tiles, itl = tile - s * i_tiles;
      const int64_t i0 = (int64_t)itl * BN, base = (int64_t)s * a.B * a.out;
      const int64_t iw = i0 + wrow;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int stage = it % NS;
        if (it >= NS) mbar_wait(smem_u32(&ctl.empty[stage]), (uint32_t)(((it / NS) - 1) & 1));
        uint8_t *As = tiles + stage * kStage, *Bs = As + 2 * A_TILE;
        const int64_t oc = (int64_t)kb * BK + chunk * 4;
        const bool col_ok = oc < a.out;
        float4 zv[4], vv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int b = wrow + 32 * j;
          const bool ok = col_ok && b < a.B;
          const int64_t idx = ok ? base + (int64_t)b * a.out + oc : 0;
          const float4 z = __ldg(reinterpret_cast<const float4 *>(a.dy + idx)), v = __ldg(reinterpret_cast<const float4 *>(dv + idx));
          zv[j] = ok ? z : make_float4(0.f, 0.f, 0.f, 0.f);
          vv[j] = ok ? v : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f), v2 = m;
        if (col_ok && iw < a.in) {
          m = __ldg(reinterpret_cast<const float4 *>(a.w_mu + iw * a.out + oc));
          const float4 r = __ldg(reinterpret_cast<const float4 *>(a.w_rho + iw * a.out + oc));
          const float s0 = softplus_fast(r.x), s1 = softplus_fast(r.y), s2 = softplus_fast(r.z), s3 = softplus_fast(r.w);
          v2 = make_float4(s0 * s0, s1 * s1, s2 * s2, s3 * s3);
        }
        *reinterpret_cast<float4 *>(Bs + t_off) = m;                  // [32 i rows][32 o]: K-major as it lies in memory
        *reinterpret_cast<float4 *>(Bs + B_TILE + t_off) = v2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          *reinterpret_cast<float4 *>(As + t_off + 4096 * j) = zv[j];
          *reinterpret_cast<float4 *>(As + A_TILE + t_off + 4096 * j) = vv[j];
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&ctl.full[stage]));
      }
      mbar_wait_parked(smem_u32(&ctl.acc), (uint32_t)(seg & 1));
      tc_fence_after_sync();
      {  // drain: lane = batch row, 16 columns per warp half; combine with x and add the partial sum into dx
        const int b = (warp & 3) * 32 + lane, half = warp >> 2;
        const float *xr = a.x + (int64_t)s * a.x_sstride + (int64_t)b * a.in + i0;
        float *dr = a.dx + ((int64_t)s * a.B + b) * a.in + i0;
#pragma unroll
        for (int c0 = 0; c0 < 16; c0 += 8) {
          const int col = half * 16 + c0;
          float p1[8], p2[8];
          tmem_ld8(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col, p1);
          tmem_ld8(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(32 + col), p2);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = col + 4 * h;
            if (b < a.B && i0 + c < a.in) {
              float4 xv = __ldg(reinterpret_cast<const float4 *>(xr + c));
              if (relu) xv = relu4(xv);
              float4 d = make_float4(dsc * fmaf(2.0f * xv.x, p2[4 * h], p1[4 * h]), dsc * fmaf(2.0f * xv.y, p2[4 * h + 1], p1[4 * h + 1]),
                                     dsc * fmaf(2.0f * xv.z, p2[4 * h + 2], p1[4 * h + 2]), dsc * fmaf(2.0f * xv.w, p2[4 * h + 3], p1[4 * h + 3]));
              if (dx_preact) d = mask4(d, xv);
              red_add_v4(dr + c, d);
            }
          }
        }
      }
      tc_fence_before_sync();
      bar_producers();
      u += kb1 - kb0;
    }
  }
  pdl_launch_dependents();
  tc_fence_before_sync();
  __syncthreads();
  if (warp == NPW) tmem_dealloc(tmem, kTmemCols);
}

// ---- wgrad ------------------------------------------------------------------------------------------------
constexpr int WG_N = 96;                                        // o window (3 regions)
constexpr int WG_CH = 32 * 128;                                 // bytes of one [32 rows][128 B] region
constexpr int WG_STAGE = (4 + 4 + 3 + 3) * WG_CH;               // x | x^2 | dz | dV
constexpr int WG_DYN = 2 * WG_STAGE + 1024;

__global__ void __launch_bounds__(256, 1) lr_wgrad_tc_kernel(const LrArgs a_in, const float *dv, int T_i) {
  extern __shared__ uint8_t dsm[];
  __shared__ Ctl ctl;
  LrArgs a = a_in;
  uint8_t *tiles = align1024(dsm);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool relu = a.flags & BBB_F_RELU_IN;
  const int64_t i_t0 = (int64_t)blockIdx.y * T_i;
  const int rows_i = (int)min((int64_t)T_i, a.in - i_t0);
  const int nq_o = (int)(a.out >> 2);
  const int q_lo = (int)((int64_t)blockIdx.x * nq_o / gridDim.x), q_hi = (int)((int64_t)(blockIdx.x + 1) * nq_o / gridDim.x);
  const int tq = q_hi - q_lo;
  const int64_t o_lo = (int64_t)q_lo * 4;
  ctl_setup(ctl, 256);
  const uint32_t tmem = ctl.tmem_base;
  pdl_wait();
  constexpr uint32_t idesc = idesc_tf32_major(128, WG_N, 1, 1);
  const int cps = (int)((a.B + 31) >> 5), nchunk = a.S * cps;

  for (int c = 0; c < nchunk; ++c) {
    const int stage = c & 1;
    if (c >= 2) mbar_wait(smem_u32(&ctl.bar[stage]), (uint32_t)(((c >> 1) - 1) & 1));
    uint8_t *X = tiles + stage * WG_STAGE, *X2 = X + 4 * WG_CH, *DZ = X2 + 4 * WG_CH, *DV = DZ + 3 * WG_CH;
    const int s = c / cps;
    const int64_t b0 = (int64_t)(c - s * cps) * 32;
    const float *xs = a.x + (int64_t)s * a.x_sstride;
    const int64_t zb = (int64_t)s * a.B * a.out;
    float4 xv[4], zv[3], vv[3];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = tid + 256 * j, r = idx >> 5, c4 = idx & 31;
      const bool ok = b0 + r < a.B && c4 * 4 < rows_i;
      const float4 v = __ldg(reinterpret_cast<const float4 *>(ok ? xs + (b0 + r) * a.in + i_t0 + c4 * 4 : a.x));
      xv[j] = ok ? v : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int idx = tid + 256 * j, r = idx / 24, c4 = idx - r * 24;
      const bool ok = b0 + r < a.B && c4 < tq;
      const int64_t off = ok ? zb + (b0 + r) * a.out + o_lo + c4 * 4 : 0;
      const float4 z = __ldg(reinterpret_cast<const float4 *>(a.dy + off)), v = __ldg(reinterpret_cast<const float4 *>(dv + off));
      zv[j] = ok ? z : make_float4(0.f, 0.f, 0.f, 0.f);
      vv[j] = ok ? v : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = tid + 256 * j, r = idx >> 5, c4 = idx & 31;
      const float4 v = relu ? relu4(xv[j]) : xv[j];
      const uint32_t off = (c4 >> 3) * WG_CH + mn32_off(r, c4 & 7);
      *reinterpret_cast<float4 *>(X + off) = v;
      *reinterpret_cast<float4 *>(X2 + off) = make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int idx = tid + 256 * j, r = idx / 24, c4 = idx - r * 24;
      const uint32_t off = (c4 >> 3) * WG_CH + mn32_off(r, c4 & 7);
      *reinterpret_cast<float4 *>(DZ + off) = zv[j];
      *reinterpret_cast<float4 *>(DV + off) = vv[j];
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after_sync();
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t acc = (c == 0 && kk == 0) ? 0u : 1u;
        mma_tf32(tmem, smem_desc_mn32(smem_u32(X) + kk * 1024, WG_CH), smem_desc_mn32(smem_u32(DZ) + kk * 1024, WG_CH), idesc, acc);
        mma_tf32(tmem + 128, smem_desc_mn32(smem_u32(X2) + kk * 1024, WG_CH), smem_desc_mn32(smem_u32(DV) + kk * 1024, WG_CH), idesc, acc);
      }
      mma_commit(smem_u32(&ctl.bar[stage]));
    }
  }
  if (tid == 0) mma_commit(smem_u32(&ctl.bar[2]));
  mbar_wait_parked(smem_u32(&ctl.bar[2]), 0);
  tc_fence_after_sync();
  pdl_launch_dependents();

  // epilogue: lane = weight row i, 48 columns per warp half; [in, out] rows are contiguous along o
  {
    const bool accum = a.flags & BBB_F_ACCUM;
    const float osc = a.out_scale_dev ? __ldg(a.out_scale_dev) : 1.0f;
    const float gk = (a.flags & BBB_F_LOGPROB) ? a.g_kl * (a.g_kl_dev ? __ldg(a.g_kl_dev) : 1.0f) : 0.0f;
    const float inv_sp2 = 1.0f / (a.sigma_p * a.sigma_p);
    const int i_r = (warp & 3) * 32 + lane, half = warp >> 2;
#pragma unroll 1
    for (int c0 = 0; c0 < 48; c0 += 8) {
      const int col = half * 48 + c0;
      float g1[8], g2[8];
      tmem_ld8(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col, g1);
      tmem_ld8(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(128 + col), g2);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c4 = (col >> 2) + h;
        if (i_r < rows_i && c4 < tq) {
          const int64_t e = (i_t0 + i_r) * a.out + o_lo + c4 * 4;
          const float4 m = __ldg(reinterpret_cast<const float4 *>(a.w_mu + e)), r = __ldg(reinterpret_cast<const float4 *>(a.w_rho + e));
          const float mu[4] = {m.x, m.y, m.z, m.w}, rho[4] = {r.x, r.y, r.z, r.w};
          float gm[4], gr[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float sg, sgm;
            softplus_sigmoid_fast(rho[j], sg, sgm);
            gm[j] = osc * fmaf(gk * mu[j], inv_sp2, g1[4 * h + j]);
            gr[j] = osc * sgm * (2.0f * sg * g2[4 * h + j] + gk * (sg * inv_sp2 - __fdividef(1.0f, sg)));
          }
          float4 *pm = reinterpret_cast<float4 *>(a.g_w_mu + e), *pr = reinterpret_cast<float4 *>(a.g_w_rho + e);
          if (accum) {
            const float4 om = *pm, orr = *pr;
            gm[0] += om.x; gm[1] += om.y; gm[2] += om.z; gm[3] += om.w;
            gr[0] += orr.x; gr[1] += orr.y; gr[2] += orr.z; gr[3] += orr.w;
          }
          *pm = make_float4(gm[0], gm[1], gm[2], gm[3]);
          *pr = make_float4(gr[0], gr[1], gr[2], gr[3]);
        }
      }
    }
  }
  ctl_teardown(ctl, 256);
}

inline int cdiv_i(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace

bool lr_tc_supported(const LrArgs &a) {
  const bool sample = a.flags & BBB_F_SAMPLE;
  return a.vec_in && a.B >= 1 && a.B <= BM && a.S >= 1 && a.in >= 4 && a.out >= 1 && (!sample || a.delta) &&
         a.S * a.B * a.out < (int64_t)1 << 31;
}

int launch_lr_fwd_tc(const LrArgs &a, cudaStream_t st) {
  const bool sample = a.flags & BBB_F_SAMPLE;
  const int o_tiles = cdiv_i(a.out, BN), nkb = cdiv_i(a.in, BK);
  const int total = (int)a.S * o_tiles * nkb;
  const int grid = total < kCtaPerSm * sm_count() ? total : kCtaPerSm * sm_count();
  if (!(a.flags & BBB_F_OUT_ZEROED)) {
    const size_t bytes = sizeof(float) * (size_t)a.S * a.B * a.out;
    BBB_CHECK_CUDA(cudaMemsetAsync(a.y, 0, bytes, st));
    if (sample) BBB_CHECK_CUDA(cudaMemsetAsync(a.delta, 0, bytes, st));
    note_launch(sample ? 2 : 1);
  }
  const int nquads = (int)((a.S * a.B * a.out) >> 2);
  if (sample) {
    BBB_CHECK_CUDA(cudaFuncSetAttribute(lr_fwd_sk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDyn));
    BBB_CHECK_CUDA(launch_pdl(lr_fwd_sk_kernel<true>, dim3(grid), dim3(NT2), kDyn, st, a, nkb, o_tiles, total));
    BBB_CHECK_LAUNCH();
    if (a.vec_out) BBB_CHECK_CUDA(launch_pdl(lr_epilogue_kernel, dim3(cdiv_i(nquads, 256)), dim3(256), 0, st, a));
    else BBB_CHECK_CUDA(launch_pdl(lr_epilogue_scalar_kernel, dim3(cdiv_i(a.S * a.B * a.out, 256)), dim3(256), 0, st, a));
  } else {   // eval with mean weights: y = x mu + mu_b, complete after the contraction
    BBB_CHECK_CUDA(cudaFuncSetAttribute(lr_fwd_sk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDyn));
    BBB_CHECK_CUDA(launch_pdl(lr_fwd_sk_kernel<false>, dim3(grid), dim3(NT2), kDyn, st, a, nkb, o_tiles, total));
  }
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

bool lr_bwd_tc_supported(const LrArgs &a) {
  return (a.flags & BBB_F_SAMPLE) && a.delta_in && !a.mask && a.vec_in && a.vec_out && a.B >= 1 && a.B <= BM && a.S >= 1 &&
         a.in >= 4 && a.out >= 4;
}

// NOTE: overwrites the forward's delta with dV (see lr_dv_kernel).
int launch_lr_bwd_tc(const LrArgs &a, cudaStream_t st) {
  float *dv = const_cast<float *>(a.delta_in);
  BBB_CHECK_CUDA(launch_pdl(lr_dv_kernel, dim3(cdiv_i(a.out, 32)), dim3(1024), 0, st, a, dv));
  BBB_CHECK_LAUNCH();
  if (!(a.flags & BBB_F_NO_DX)) {
    if (!(a.flags & BBB_F_OUT_ZEROED)) {
      BBB_CHECK_CUDA(cudaMemsetAsync(a.dx, 0, sizeof(float) * (size_t)a.S * a.B * a.in, st));
      note_launch();
    }
    const int i_tiles = cdiv_i(a.in, BN), nkb = cdiv_i(a.out, BK);
    const int total = (int)a.S * i_tiles * nkb;
    const int grid = total < kCtaPerSm * sm_count() ? total : kCtaPerSm * sm_count();
    BBB_CHECK_CUDA(cudaFuncSetAttribute(lr_dgrad_sk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDyn));
    BBB_CHECK_CUDA(launch_pdl(lr_dgrad_sk_kernel, dim3(grid), dim3(NT2), kDyn, st, a, (const float *)dv, nkb, i_tiles, total));
    BBB_CHECK_LAUNCH();
  }
  if (a.flags & BBB_F_NO_WGRAD) return BBB_OK;
  const int n_it = cdiv_i(a.in, 128);
  const int T_i = ((cdiv_i(a.in, n_it) + 3) / 4) * 4;
  const int nq_o = (int)(a.out / 4);
  int n_c = sm_count() / n_it;
  const int need = cdiv_i(nq_o, (WG_N - 8) / 4);
  if (n_c < need) n_c = need;
  if (n_c > nq_o) n_c = nq_o;
  if (n_c < 1) n_c = 1;
  BBB_CHECK_CUDA(cudaFuncSetAttribute(lr_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_DYN));
  BBB_CHECK_CUDA(launch_pdl(lr_wgrad_tc_kernel, dim3(n_c, cdiv_i(a.in, T_i)), dim3(256), WG_DYN, st, a, (const float *)dv, T_i));
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

}  // namespace bbb
