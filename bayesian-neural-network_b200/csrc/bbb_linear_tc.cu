// Weight-sampling Bayesian linear layer on the 5th-gen tensor cores (tcgen05, kind::tf32, TMEM).
//
// Design (DESIGN.md "tcgen05 path"): weight-stationary tiles.  A CTA owns a [BN x k_chunk] tile of the
// weight matrix and ALL batch rows of one 128-row M tile.  For every 32-wide k block its 256 threads
//   * read mu/rho once (16-byte vector loads), form sigma once, and for each of the SG Monte-Carlo samples
//     handled by the CTA generate eps (Philox, or injected), W = mu + sigma eps, and the log-prob terms,
//   * write the W_s tiles and the activation tile straight into shared memory in the canonical K-major
//     SWIZZLE_128B operand layout (W is never written to global memory),
//   * fence the generic-proxy writes to the async proxy, and one elected thread issues
//     tcgen05.mma.cta_group::1.kind::tf32 (M = 128, N = BN, K = 8) x 4 x SG into SG TMEM accumulators.
// Two shared-memory stages let the tensor pipe run block k while the threads stage block k+1
// (tcgen05.commit -> mbarrier releases a stage).  The epilogue drains TMEM with tcgen05.ld.
//   fwd   : y_s[b][o]  (split-K partial sums are combined with fp32 red.global.add when the K range is split)
//   dgrad : dx_s[b][i] = sum_o dz_s[b][o] W_s[o][i]   (W tile transposed in registers while staging)
//   wgrad : G_s[i][o]  = sum_b x_s[b][i] dz_s[b][o], then the analytic mu/rho gradient epilogue with eps
//           regenerated, summed over samples.
// mu/rho traffic per call: read once per CTA tile = 8 B/weight for ALL samples of the group.
#include "bbb_tc_tiles.cuh"

namespace bbb {
namespace {

using namespace tc;
using namespace tcx;

// ==================================================================================================
// forward
// ==================================================================================================
template <int BN, int SG, bool kLogProb>
__global__ void __launch_bounds__(NT, 2) fwd_tc_kernel(const LinArgs a_in, int k_chunk, int n_ksplit) {
  using SM = Smem<BN, SG>;
  extern __shared__ uint8_t dsm[];
  __shared__ Ctl ctl;
  __shared__ float bias_s[SG][BN];
  __shared__ float red[64];
  LinArgs a = a_in;
  rng_resolve(a.rng);
  uint8_t *tiles = align1024(dsm);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t o0 = (int64_t)blockIdx.x * BN, m0 = (int64_t)blockIdx.y * BM;
  const int ksplit = blockIdx.z % n_ksplit, sgroup = blockIdx.z / n_ksplit;
  const int s0 = sgroup * SG, ns = min(SG, a.S - s0);
  const bool sample = a.flags & BBB_F_SAMPLE, relu = a.flags & BBB_F_RELU_IN, x_shared = a.x_sstride == 0;
  const bool lpcta = kLogProb && blockIdx.y == 0;
  const int64_t k_begin = (int64_t)ksplit * k_chunk, k_end = min(a.in, k_begin + (int64_t)k_chunk);
  const int nkb = k_end > k_begin ? (int)((k_end - k_begin + BK - 1) / BK) : 0;
  constexpr uint32_t idesc = idesc_tf32(BM, BN);

  ctl_setup(ctl, SM::kTmemCols);
  const uint32_t tmem = ctl.tmem_base;

  float lp[SG], lq[SG];
#pragma unroll
  for (int s = 0; s < SG; ++s) lp[s] = lq[s] = 0.0f;

  if (tid < BN) {
    const int64_t o = o0 + tid;
#pragma unroll
    for (int s = 0; s < SG; ++s) {
      float bv = 0.0f;
      if (s < ns && o < a.out) {
        float sg, ep;
        bias_elem(a, s0 + s, o, sample, lpcta, bv, sg, ep);
        if (lpcta && ksplit == 0) { lp[s] += logp_elem(a.prior, bv); lq[s] += logq_elem(sg, ep); }
      }
      bias_s[s][tid] = bv;
    }
  }

  for (int it = 0; it < nkb; ++it) {
    const int stage = it & 1;
    if (it >= 2) mbar_wait(smem_u32(&ctl.bar[stage]), (uint32_t)(((it >> 1) - 1) & 1));
    uint8_t *As = tiles + stage * SM::kStage, *Bs = As + SG * A_TILE;
    const int64_t kb = k_begin + (int64_t)it * BK;
    // activations: [128 batch rows][32 k], one tile per sample unless the input is shared
#pragma unroll
    for (int s = 0; s < SG; ++s) {
      if (s < ns && (s == 0 || !x_shared)) {
        const float *xs = a.x + (int64_t)(s0 + s) * a.x_sstride;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int idx = tid + NT * j, row = idx >> 3, chunk = idx & 7;
          const int64_t k = kb + chunk * 4;
          float4 v = k < k_end ? ld_row4(xs, m0 + row, k, a.B, a.in, true) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (relu) v = relu4(v);
          st_tile4(As + s * A_TILE, row, chunk, v.x, v.y, v.z, v.w);
        }
      }
    }
    // weights: [BN out rows][32 k] per sample, formed in registers
    for (int idx = tid; idx < BN * 8; idx += NT) {
      const int row = idx >> 3, chunk = idx & 7;
      const int64_t o = o0 + row, k = kb + chunk * 4;
      if (o < a.out && k < k_end) {
        const int64_t e = o * a.in + k;
        Quad q;
        load_quad(a, e, sample || lpcta, q);
        float lsg = 0.0f;
        if (lpcta) lsg = __logf(q.sg[0]) + __logf(q.sg[1]) + __logf(q.sg[2]) + __logf(q.sg[3]);
#pragma unroll
        for (int s = 0; s < SG; ++s) {
          if (s < ns) {
            float ep[4], w[4];
            sample_quad(a, s0 + s, e, q, sample, ep, w);
            st_tile4(Bs + s * SM::kB, row, chunk, w[0], w[1], w[2], w[3]);
            if (lpcta) {
              lp[s] += logp_elem_fast(a.prior, w[0]) + logp_elem_fast(a.prior, w[1]) +
                       logp_elem_fast(a.prior, w[2]) + logp_elem_fast(a.prior, w[3]);
              lq[s] += -4.0f * kHalfLog2Pi - lsg -
                       0.5f * (ep[0] * ep[0] + ep[1] * ep[1] + ep[2] * ep[2] + ep[3] * ep[3]);
            }
          }
        }
      } else {
#pragma unroll
        for (int s = 0; s < SG; ++s)
          if (s < ns) st_tile4(Bs + s * SM::kB, row, chunk, 0.f, 0.f, 0.f, 0.f);
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after_sync();
#pragma unroll
      for (int s = 0; s < SG; ++s)
        if (s < ns)
          issue_block(tmem + s * BN, smem_u32(As + (x_shared ? 0 : s) * A_TILE), smem_u32(Bs + s * SM::kB), idesc,
                      it == 0);
      mma_commit(smem_u32(&ctl.bar[stage]));
    }
  }
  if (tid == 0) mma_commit(smem_u32(&ctl.bar[2]));
  mbar_wait(smem_u32(&ctl.bar[2]), 0);
  tc_fence_after_sync();
  __syncthreads();  // bias_s

  // epilogue: TMEM -> swizzled smem -> coalesced 16-byte stores / red.v4 (split-K)
  {
    const bool add_bias = ksplit == 0;
    drain_tile<BN, SG>(tiles, tmem, ns, nkb > 0, a.y + (int64_t)s0 * a.B * a.out, a.B * a.out, m0, a.B, o0, a.out,
                       a.vec_out, n_ksplit > 1, 1.0f,
                       [&](int s, int c) { return add_bias ? bias_s[s][c] : 0.0f; });
  }
  if (lpcta) {
#pragma unroll
    for (int s = 0; s < SG; ++s)
      if (s < ns) block_sum2_atomic(lp[s], lq[s], red, a.logp + s0 + s, a.logq + s0 + s);
  }
  ctl_teardown(ctl, SM::kTmemCols);
}

// ==================================================================================================
// dgrad: dx_s[b][i] = sum_o dz_s[b][o] W_s[o][i];  A = dz (K = o, natural), B[i][o] = W^T (register transpose)
// ==================================================================================================
template <int BN, int SG>
__global__ void __launch_bounds__(NT, 2) dgrad_tc_kernel(const LinArgs a_in, int k_chunk, int n_ksplit) {
  using SM = Smem<BN, SG>;
  extern __shared__ uint8_t dsm[];
  __shared__ Ctl ctl;
  LinArgs a = a_in;
  rng_resolve(a.rng);
  uint8_t *tiles = align1024(dsm);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t i0 = (int64_t)blockIdx.x * BN, m0 = (int64_t)blockIdx.y * BM;
  const int ksplit = blockIdx.z % n_ksplit, sgroup = blockIdx.z / n_ksplit;
  const int s0 = sgroup * SG, ns = min(SG, a.S - s0);
  const bool sample = a.flags & BBB_F_SAMPLE;
  const int64_t k_begin = (int64_t)ksplit * k_chunk, k_end = min(a.out, k_begin + (int64_t)k_chunk);
  const int nkb = k_end > k_begin ? (int)((k_end - k_begin + BK - 1) / BK) : 0;
  constexpr uint32_t idesc = idesc_tf32(BM, BN);

  ctl_setup(ctl, SM::kTmemCols);
  const uint32_t tmem = ctl.tmem_base;

  for (int it = 0; it < nkb; ++it) {
    const int stage = it & 1;
    if (it >= 2) mbar_wait(smem_u32(&ctl.bar[stage]), (uint32_t)(((it >> 1) - 1) & 1));
    uint8_t *As = tiles + stage * SM::kStage, *Bs = As + SG * A_TILE;
    const int64_t kb = k_begin + (int64_t)it * BK;
    // dz tile [128 b][32 o] per sample (mask = ReLU of this layer's own output)
#pragma unroll
    for (int s = 0; s < SG; ++s) {
      if (s < ns) {
        const int64_t base = (int64_t)(s0 + s) * a.B * a.out;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int idx = tid + NT * j, row = idx >> 3, chunk = idx & 7;
          const int64_t o = kb + chunk * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (o < k_end) {
            v = ld_row4(a.dy + base, m0 + row, o, a.B, a.out, a.vec_out);
            if (a.mask) v = mask4(v, ld_row4(a.mask + base, m0 + row, o, a.B, a.out, a.vec_out));
            if (o + 3 >= k_end) {  // columns past this CTA's K range belong to the next split
              if (o + 1 >= k_end) v.y = 0.f;
              if (o + 2 >= k_end) v.z = 0.f;
              if (o + 3 >= k_end) v.w = 0.f;
            }
          }
          st_tile4(As + s * A_TILE, row, chunk, v.x, v.y, v.z, v.w);
        }
      }
    }
    // W^T tile [BN i rows][32 o]: 4x4 blocks (4 weight rows o, one i quad), transposed in registers
    for (int idx = tid; idx < 8 * (BN / 4); idx += NT) {
      const int iq = idx % (BN / 4), oq = idx / (BN / 4);
      const int64_t i = i0 + iq * 4;
      float w[SG][4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int64_t o = kb + oq * 4 + r;
        if (o < k_end && i < a.in) {
          const int64_t e = o * a.in + i;
          Quad q;
          load_quad(a, e, sample, q);
#pragma unroll
          for (int s = 0; s < SG; ++s) {
            float ep[4];
            if (s < ns) sample_quad(a, s0 + s, e, q, sample, ep, w[s][r]);
          }
        } else {
#pragma unroll
          for (int s = 0; s < SG; ++s)
#pragma unroll
            for (int c = 0; c < 4; ++c) w[s][r][c] = 0.0f;
        }
      }
#pragma unroll
      for (int s = 0; s < SG; ++s)
        if (s < ns) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            st_tile4(Bs + s * SM::kB, iq * 4 + c, oq, w[s][0][c], w[s][1][c], w[s][2][c], w[s][3][c]);
        }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after_sync();
#pragma unroll
      for (int s = 0; s < SG; ++s)
        if (s < ns) issue_block(tmem + s * BN, smem_u32(As + s * A_TILE), smem_u32(Bs + s * SM::kB), idesc, it == 0);
      mma_commit(smem_u32(&ctl.bar[stage]));
    }
  }
  if (tid == 0) mma_commit(smem_u32(&ctl.bar[2]));
  mbar_wait(smem_u32(&ctl.bar[2]), 0);
  tc_fence_after_sync();

  const float osc = ((a.flags & BBB_F_SCALE_DX) && a.out_scale_dev) ? __ldg(a.out_scale_dev) : 1.0f;
  drain_tile<BN, SG>(tiles, tmem, ns, nkb > 0, a.dx + (int64_t)s0 * a.B * a.in, a.B * a.in, m0, a.B, i0, a.in, a.vec_in,
                     n_ksplit > 1, osc, [](int, int) { return 0.0f; },
                     (a.flags & BBB_F_DX_PREACT) ? a.x + (int64_t)s0 * a.x_sstride : nullptr);
  ctl_teardown(ctl, SM::kTmemCols);
}

// ==================================================================================================
// wgrad: G_s[i][o] = sum_b x_s[b][i] dz_s[b][o]  (M = i tile of 128, N = o tile of BN, K = batch), then
//   t = G - gp w R(w);  grad_mu += t;  grad_rho += sigmoid(rho) (t eps - gq / sigma)      (eps regenerated)
// ==================================================================================================
template <int BN, int SG>
__global__ void __launch_bounds__(NT, 2) wgrad_tc_kernel(const LinArgs a_in) {
  using SM = Smem<BN, SG>;
  static_assert(SG * BN * BM * 4 <= SM::kTiles, "G staging must fit in the operand buffers");
  extern __shared__ uint8_t dsm[];
  __shared__ Ctl ctl;
  LinArgs a = a_in;
  rng_resolve(a.rng);
  uint8_t *tiles = align1024(dsm);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t o0 = (int64_t)blockIdx.x * BN, i0 = (int64_t)blockIdx.y * BM;
  const bool sample = a.flags & BBB_F_SAMPLE, relu = a.flags & BBB_F_RELU_IN;
  const bool bias_cta = blockIdx.y == 0;
  const int nkb = (int)((a.B + BK - 1) / BK);
  constexpr uint32_t idesc = idesc_tf32(BM, BN);
  const float osc = a.out_scale_dev ? __ldg(a.out_scale_dev) : 1.0f;
  const bool accum_flag = a.flags & BBB_F_ACCUM;

  ctl_setup(ctl, SM::kTmemCols);
  const uint32_t tmem = ctl.tmem_base;
  float *Gs = reinterpret_cast<float *>(tiles);  // [SG][BN][128] after the MMAs of a group have completed
  __shared__ float colsum_s[SG][BN];             // sum_b dz[b][o] (bias gradients), first i tile only

  int it_global = 0;  // pipeline iteration counter across sample groups (mbarrier phases keep running)
  const int ngroups = (a.S + SG - 1) / SG;
  for (int g = 0; g < ngroups; ++g) {
    const int s0 = g * SG, ns = min(SG, a.S - s0);
    float bsum[SG][4];
#pragma unroll
    for (int s = 0; s < SG; ++s) bsum[s][0] = bsum[s][1] = bsum[s][2] = bsum[s][3] = 0.0f;
    if (bias_cta && tid < BN) {
#pragma unroll
      for (int s = 0; s < SG; ++s) colsum_s[s][tid] = 0.0f;
    }
    for (int it = 0; it < nkb; ++it, ++it_global) {
      const int stage = it_global & 1;
      if (it >= 2) mbar_wait(smem_u32(&ctl.bar[stage]), (uint32_t)(((it_global >> 1) - 1) & 1));
      uint8_t *As = tiles + stage * SM::kStage, *Bs = As + SG * A_TILE;
      const int64_t b0 = (int64_t)it * BK;
      // x^T tile [128 i rows][32 b]: one 4x4 block per thread (8 b quads x 32 i quads)
#pragma unroll
      for (int s = 0; s < SG; ++s) {
        if (s < ns && (s == 0 || a.x_sstride != 0)) {
          const float *xs = a.x + (int64_t)(s0 + s) * a.x_sstride;
          const int iq = tid & 31, bq = tid >> 5;
          float4 v[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            v[r] = ld_row4(xs, b0 + bq * 4 + r, i0 + iq * 4, a.B, a.in, true);
            if (relu) v[r] = relu4(v[r]);
          }
          uint8_t *T = As + s * A_TILE;
          st_tile4(T, iq * 4 + 0, bq, v[0].x, v[1].x, v[2].x, v[3].x);
          st_tile4(T, iq * 4 + 1, bq, v[0].y, v[1].y, v[2].y, v[3].y);
          st_tile4(T, iq * 4 + 2, bq, v[0].z, v[1].z, v[2].z, v[3].z);
          st_tile4(T, iq * 4 + 3, bq, v[0].w, v[1].w, v[2].w, v[3].w);
        }
      }
      // dz^T tile [BN o rows][32 b]
#pragma unroll
      for (int s = 0; s < SG; ++s) {
        if (s < ns) {
          const int64_t base = (int64_t)(s0 + s) * a.B * a.out;
          for (int idx = tid; idx < 8 * (BN / 4); idx += NT) {
            const int oq = idx % (BN / 4), bq = idx / (BN / 4);
            float4 v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              v[r] = ld_row4(a.dy + base, b0 + bq * 4 + r, o0 + oq * 4, a.B, a.out, a.vec_out);
              if (a.mask) v[r] = mask4(v[r], ld_row4(a.mask + base, b0 + bq * 4 + r, o0 + oq * 4, a.B, a.out, a.vec_out));
            }
            if (bias_cta) {  // each thread owns at most one (oq, bq) block: keep its column sums in registers
              bsum[s][0] += v[0].x + v[1].x + v[2].x + v[3].x;
              bsum[s][1] += v[0].y + v[1].y + v[2].y + v[3].y;
              bsum[s][2] += v[0].z + v[1].z + v[2].z + v[3].z;
              bsum[s][3] += v[0].w + v[1].w + v[2].w + v[3].w;
            }
            uint8_t *T = Bs + s * SM::kB;
            st_tile4(T, oq * 4 + 0, bq, v[0].x, v[1].x, v[2].x, v[3].x);
            st_tile4(T, oq * 4 + 1, bq, v[0].y, v[1].y, v[2].y, v[3].y);
            st_tile4(T, oq * 4 + 2, bq, v[0].z, v[1].z, v[2].z, v[3].z);
            st_tile4(T, oq * 4 + 3, bq, v[0].w, v[1].w, v[2].w, v[3].w);
          }
        }
      }
      fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after_sync();
#pragma unroll
        for (int s = 0; s < SG; ++s)
          if (s < ns)
            issue_block(tmem + s * BN, smem_u32(As + (a.x_sstride == 0 ? 0 : s) * A_TILE), smem_u32(Bs + s * SM::kB),
                        idesc, it == 0);
        mma_commit(smem_u32(&ctl.bar[stage]));
      }
    }
    if (bias_cta && tid < 8 * (BN / 4)) {
      const int oq = tid % (BN / 4);
#pragma unroll
      for (int s = 0; s < SG; ++s)
#pragma unroll
        for (int c = 0; c < 4; ++c) atomicAdd(&colsum_s[s][oq * 4 + c], bsum[s][c]);
    }
    // all MMAs of this group done -> operand buffers are free, accumulators are final
    if (tid == 0) mma_commit(smem_u32(&ctl.bar[2]));
    mbar_wait(smem_u32(&ctl.bar[2]), (uint32_t)(g & 1));
    tc_fence_after_sync();

    // TMEM -> Gs[s][o_local][i_local]
    {
      const int q4 = warp & 3, half = warp >> 2;
      constexpr int HALF = BN / 2;
#pragma unroll
      for (int s = 0; s < SG; ++s) {
        if (s >= ns) break;
#pragma unroll
        for (int c0 = 0; c0 < HALF; c0 += 8) {
          const int col = half * HALF + c0;
          float v[8];
          if (nkb > 0) {
            tmem_ld8(tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(s * BN + col), v);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.0f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) Gs[(s * BN + col + j) * BM + q4 * 32 + lane] = v[j];
        }
      }
    }
    tc_fence_before_sync();
    __syncthreads();

    float gps[SG], gqs[SG];
#pragma unroll
    for (int s = 0; s < SG; ++s) {
      gps[s] = s < ns ? a.gp * (a.gp_dev ? __ldg(a.gp_dev + (s0 + s) * a.g_dev_stride) : 1.0f) : 0.0f;
      gqs[s] = s < ns ? a.gq * (a.gq_dev ? __ldg(a.gq_dev + (s0 + s) * a.g_dev_stride) : 1.0f) : 0.0f;
    }
    const bool accum = accum_flag || g > 0;
    // analytic epilogue over (o, i quad): coalesced 16-byte reads of mu/rho, writes of grad_mu/grad_rho
    for (int idx = tid; idx < BN * 32; idx += NT) {
      const int ol = idx >> 5, iq = idx & 31;
      const int64_t o = o0 + ol, i = i0 + iq * 4;
      if (o < a.out && i < a.in) {
        const int64_t e = o * a.in + i;
        Quad q;
        load_quad(a, e, true, q);
        float gm[4] = {0.f, 0.f, 0.f, 0.f}, gr[4] = {0.f, 0.f, 0.f, 0.f}, sgm[4], isg[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { sgm[j] = sigmoid_fast(q.rho[j]); isg[j] = __fdividef(1.0f, q.sg[j]); }
#pragma unroll
        for (int s = 0; s < SG; ++s) {
          if (s < ns) {
            float ep[4], w[4];
            sample_quad(a, s0 + s, e, q, sample, ep, w);
            const float4 G = *reinterpret_cast<const float4 *>(&Gs[(s * BN + ol) * BM + iq * 4]);
            const float Gv[4] = {G.x, G.y, G.z, G.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float t = Gv[j];
              if (gps[s] != 0.0f) t = fmaf(-gps[s] * w[j], prior_R_fast(a.prior, w[j]), t);
              gm[j] += t;
              gr[j] += sgm[j] * fmaf(t, ep[j], -gqs[s] * isg[j]);
            }
          }
        }
        float4 *pm = reinterpret_cast<float4 *>(a.g_w_mu + e), *pr = reinterpret_cast<float4 *>(a.g_w_rho + e);
        float4 om = make_float4(0.f, 0.f, 0.f, 0.f), orr = om;
        if (accum) { om = *pm; orr = *pr; }
        *pm = make_float4(fmaf(osc, gm[0], om.x), fmaf(osc, gm[1], om.y), fmaf(osc, gm[2], om.z), fmaf(osc, gm[3], om.w));
        *pr = make_float4(fmaf(osc, gr[0], orr.x), fmaf(osc, gr[1], orr.y), fmaf(osc, gr[2], orr.z), fmaf(osc, gr[3], orr.w));
      }
    }
    // bias gradients: column sums of dz, by the CTAs of the first i tile
    if (bias_cta && tid < BN) {
      const int64_t o = o0 + tid;
      if (o < a.out) {
        float gbm = 0.0f, gbr = 0.0f;
        for (int s = 0; s < ns; ++s) {
          const float colsum = colsum_s[s][tid];
          float bv, sg, ep;
          bias_elem(a, s0 + s, o, sample, true, bv, sg, ep);
          float t = colsum;
          if (gps[s] != 0.0f) t = fmaf(-gps[s] * bv, prior_R(a.prior, bv), t);
          gbm += t;
          gbr += -expm1f(-sg) * (t * ep - gqs[s] / sg);
        }
        a.g_b_mu[o] = accum ? fmaf(osc, gbm, a.g_b_mu[o]) : osc * gbm;
        a.g_b_rho[o] = accum ? fmaf(osc, gbr, a.g_b_rho[o]) : osc * gbr;
      }
    }
    __syncthreads();  // Gs (operand buffers) is reused by the next group
  }
  ctl_teardown(ctl, SM::kTmemCols);
}

// ==================================================================================================
// host-side launch logic
// ==================================================================================================
inline int cdiv_i(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// K split so that the grid is about one wave of the 148 SMs; chunk is a multiple of 32
inline void pick_ksplit(int64_t K, int tiles, int *k_chunk, int *n_ksplit) {
  const int nkb = cdiv_i(K, BK);
  int want = tiles >= sm_count() ? 1 : (sm_count() + tiles / 2) / tiles;  // about one CTA per SM: every split costs a red pass
  if (want > nkb) want = nkb;
  if (want < 1) want = 1;
  const int per = cdiv_i(nkb, want);
  *k_chunk = per * BK;
  *n_ksplit = cdiv_i(nkb, per);
}
inline int pick_bn(int64_t n) { return n <= 16 ? 16 : (n % 80 == 0 ? 80 : 64); }

template <class K>
int set_smem(K kernel, int bytes) {
  BBB_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return BBB_OK;
}

template <int BN, int SG>
int launch_fwd(const LinArgs &a, cudaStream_t st) {
  const int o_tiles = cdiv_i(a.out, BN), m_tiles = cdiv_i(a.B, BM), s_groups = cdiv_i(a.S, SG);
  int k_chunk, n_ksplit;
  pick_ksplit(a.in, o_tiles * m_tiles * s_groups, &k_chunk, &n_ksplit);
  if (n_ksplit > 1) {
    BBB_CHECK_CUDA(cudaMemsetAsync(a.y, 0, sizeof(float) * (size_t)a.S * a.B * a.out, st));
    note_launch();
  }
  dim3 grid(o_tiles, m_tiles, n_ksplit * s_groups);
  const int smem = Smem<BN, SG>::kDyn;
  if (a.flags & BBB_F_LOGPROB) {
    if (int r = set_smem(fwd_tc_kernel<BN, SG, true>, smem)) return r;
    fwd_tc_kernel<BN, SG, true><<<grid, NT, smem, st>>>(a, k_chunk, n_ksplit);
  } else {
    if (int r = set_smem(fwd_tc_kernel<BN, SG, false>, smem)) return r;
    fwd_tc_kernel<BN, SG, false><<<grid, NT, smem, st>>>(a, k_chunk, n_ksplit);
  }
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

template <int BN, int SG>
int launch_dgrad(const LinArgs &a, cudaStream_t st) {
  const int i_tiles = cdiv_i(a.in, BN), m_tiles = cdiv_i(a.B, BM), s_groups = cdiv_i(a.S, SG);
  int k_chunk, n_ksplit;
  pick_ksplit(a.out, i_tiles * m_tiles * s_groups, &k_chunk, &n_ksplit);
  if (n_ksplit > 1) {
    BBB_CHECK_CUDA(cudaMemsetAsync(a.dx, 0, sizeof(float) * (size_t)a.S * a.B * a.in, st));
    note_launch();
  }
  dim3 grid(i_tiles, m_tiles, n_ksplit * s_groups);
  const int smem = Smem<BN, SG>::kDyn;
  if (int r = set_smem(dgrad_tc_kernel<BN, SG>, smem)) return r;
  dgrad_tc_kernel<BN, SG><<<grid, NT, smem, st>>>(a, k_chunk, n_ksplit);
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

template <int BN, int SG>
int launch_wgrad(const LinArgs &a, cudaStream_t st) {
  dim3 grid(cdiv_i(a.out, BN), cdiv_i(a.in, BM));
  const int smem = Smem<BN, SG>::kDyn;
  if (int r = set_smem(wgrad_tc_kernel<BN, SG>, smem)) return r;
  wgrad_tc_kernel<BN, SG><<<grid, NT, smem, st>>>(a);
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

#define BBB_DISPATCH_BN_SG(FN, bn, sg, ...)                   \
  do {                                                        \
    if (sg == 1) {                                            \
      if (bn == 16) return FN<16, 1>(__VA_ARGS__);            \
      if (bn == 80) return FN<80, 1>(__VA_ARGS__);            \
      return FN<64, 1>(__VA_ARGS__);                          \
    } else {                                                  \
      if (bn == 16) return FN<16, 2>(__VA_ARGS__);            \
      if (bn == 80) return FN<80, 2>(__VA_ARGS__);            \
      return FN<64, 2>(__VA_ARGS__);                          \
    }                                                         \
  } while (0)

}  // namespace

bool linear_tc_supported(const LinArgs &a) {
  // 16-byte rows of mu/rho/x/eps (vector loads, aligned Philox quads); anything else takes the FMA path
  return a.vec_in && a.in >= 4 && a.out >= 1 && a.B >= 1 && a.S >= 1;
}

int launch_linear_fwd_tc(const LinArgs &a, cudaStream_t st) {
  const int bn = pick_bn(a.out), sg = a.S >= 2 ? 2 : 1;
  BBB_DISPATCH_BN_SG(launch_fwd, bn, sg, a, st);
}

int launch_linear_bwd_tc(const LinArgs &a, cudaStream_t st) {
  const int sg = a.S >= 2 ? 2 : 1;
  if (!(a.flags & BBB_F_NO_DX)) {
    const int bn = pick_bn(a.in);
    int r = [&]() -> int { BBB_DISPATCH_BN_SG(launch_dgrad, bn, sg, a, st); }();
    if (r) return r;
  }
  if (a.flags & BBB_F_NO_WGRAD) return BBB_OK;
  const int bn = pick_bn(a.out);
  BBB_DISPATCH_BN_SG(launch_wgrad, bn, sg, a, st);
}

}  // namespace bbb
