// Weight-sampling Bayesian linear layer, tcgen05 kind::tf32, for batches of at most 128 rows (one UMMA M tile):
// the regime of BASELINE.json configs 1-4, where the contraction is tiny and the kernel is bound by the
// sampling work per weight (Philox + Box-Muller + prior/posterior terms), not by the tensor pipe.
//
// What decides the speed here is how evenly that per-weight work is spread over the 148 SMs and whether every
// SM has enough warps in flight to issue it, so the three kernels are organised around that:
//
//   fwd / dgrad  stream-K.  The weight matrix is cut into units of [BN rows x 32 k]; the units, in (tile, k)
//                order, are dealt out in equal contiguous ranges to 2 x 148 CTAs (two co-resident CTAs per SM,
//                18 warps).  A CTA accumulates each run of units that belongs to one output tile in TMEM and
//                adds that partial tile to the (zero-filled) output with red.global.add.v4.f32.
//                Warp-specialised: 8 producer warps sample W_s = mu + sigma eps into the SWIZZLE_128B operand
//                ring (full/empty mbarriers, no CTA-wide barrier in the loop), 1 warp issues tcgen05.mma.
//   wgrad        balanced slabs.  The input dimension is cut into equal row tiles of T <= 128 rows and each row
//                tile's output columns are split evenly (granularity 1) over its share of the 2 x 148 CTAs.  A CTA
//                runs G_s = x_s^T dz_s for a 16-aligned window that covers its columns (K = batch), then the
//                analytic mu/rho-gradient epilogue with eps regenerated, for exactly its own weights.
// mu and rho are read once per launch for all samples of a group; W never leaves the SM.
#include "bbb_tc_tiles.cuh"

namespace bbb {
namespace {

using namespace tc;
using namespace tcx;

constexpr int NPW = 8;            // producer warps
constexpr int NTP = NPW * 32;     // producer threads (the drain helpers assume 256)
constexpr int NT2 = NTP + 32;     // + the MMA-issuing warp
constexpr int NS = 2;             // operand ring depth
constexpr int kCtaPerSm = 2;

struct Ctl2 {
  uint64_t full[NS], empty[NS], acc;
  uint32_t tmem_base;
};

template <int BN, int SG>
struct Ring {
  static constexpr int kB = BN * 128;
  static constexpr int kStage = SG * A_TILE + SG * kB;
  static constexpr int kTiles = NS * kStage;
  static constexpr int kDyn = kTiles + 1024;
  static constexpr uint32_t kTmemCols = tmem_cols_pow2(SG * BN);
  static_assert(SG * BM * (((BN / 4 + 7) / 8) * 8) * 16 <= kTiles, "drain staging must fit in the ring");
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_producers() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void ctl2_setup(Ctl2 &c, uint32_t tmem_cols) {
  const int warp = threadIdx.x >> 5;
  if (warp == NPW) tmem_alloc(smem_u32(&c.tmem_base), tmem_cols);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      mbar_init(smem_u32(&c.full[s]), NPW);  // one arrival per producer warp
      mbar_init(smem_u32(&c.empty[s]), 1);   // tcgen05.commit
    }
    mbar_init(smem_u32(&c.acc), 1);
    mbar_fence_init();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
}
__device__ __forceinline__ void ctl2_teardown(Ctl2 &c, uint32_t tmem_cols) {
  tc_fence_before_sync();
  __syncthreads();
  if ((threadIdx.x >> 5) == NPW) tmem_dealloc(c.tmem_base, tmem_cols);
}

// this CTA's contiguous range of stream-K units
__device__ __forceinline__ void unit_range(int total, int &u0, int &u1) {
  u0 = (int)((int64_t)blockIdx.x * total / gridDim.x);
  u1 = (int)((int64_t)(blockIdx.x + 1) * total / gridDim.x);
}

// The MMA warp: for every unit wait for the producers, issue 4 x ns MMAs, release the stage; after the last
// unit of a tile segment signal the accumulators.  One lane issues; the warp reconverges before returning.
template <int BN, int SG>
__device__ __forceinline__ void mma_warp(Ctl2 &ctl, uint8_t *tiles, uint32_t tmem, int u0, int u1, int nkb, int n_tiles,
                                         int S, bool a_shared) {
  using R = Ring<BN, SG>;
  constexpr uint32_t idesc = idesc_tf32(BM, BN);
  if ((threadIdx.x & 31) == 0) {
    int it = 0;
    for (int u = u0; u < u1;) {
      const int tile = u / nkb, kb0 = u - tile * nkb, kb1 = min(nkb, kb0 + (u1 - u));
      const int sgi = tile / n_tiles, ns = min(SG, S - sgi * SG);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int stage = it % NS;
        mbar_wait_parked(smem_u32(&ctl.full[stage]), (uint32_t)((it / NS) & 1));
        tc_fence_after_sync();
        uint8_t *As = tiles + stage * R::kStage, *Bs = As + SG * A_TILE;
#pragma unroll
        for (int s = 0; s < SG; ++s)
          if (s < ns)
            issue_block(tmem + s * BN, smem_u32(As + (a_shared ? 0 : s) * A_TILE), smem_u32(Bs + s * R::kB), idesc,
                        kb == kb0);
        mma_commit(smem_u32(&ctl.empty[stage]));
      }
      mma_commit(smem_u32(&ctl.acc));
      u += kb1 - kb0;
    }
  }
  __syncwarp();
}

// ==================================================================================================
// forward: y_s[b][o] += sum_{k in segment} x_s[b][k] W_s[o][k]   (+ b_s[o] from the segment that starts at k = 0)
// ==================================================================================================
template <int BN, int SG, bool kLogProb>
__global__ void __launch_bounds__(NT2, kCtaPerSm) fwd_sk_kernel(const LinArgs a_in, int nkb, int o_tiles, int total) {
  using R = Ring<BN, SG>;
  extern __shared__ uint8_t dsm[];
  __shared__ Ctl2 ctl;
  __shared__ float bias_s[SG][BN];
  __shared__ float red[64];
  LinArgs a = a_in;
  rng_resolve(a.rng);
  uint8_t *tiles = align1024(dsm);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool sample = a.flags & BBB_F_SAMPLE, relu = a.flags & BBB_F_RELU_IN, x_shared = a.x_sstride == 0;
  int u0, u1;
  unit_range(total, u0, u1);

  ctl2_setup(ctl, R::kTmemCols);
  const uint32_t tmem = ctl.tmem_base;

  float lp[SG], lq[SG];
#pragma unroll
  for (int s = 0; s < SG; ++s) lp[s] = lq[s] = 0.0f;
  // thread t owns row t/8, 16-byte chunk t%8 of the weight tile and rows t/8 + 32 j of the activation tile; in the
  // SWIZZLE_128B layout row r and row r + 32 share (r & 7), so one offset (+ 4096 j) serves all of them
  const int wrow = tid >> 3, chunk = tid & 7;
  const uint32_t w_off = sw128_off(wrow, chunk);
  int xrow_off[4];
  bool xrow_ok[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    xrow_ok[j] = wrow + 32 * j < a.B;
    xrow_off[j] = (wrow + 32 * j) * (int)a.in;
  }

  if (warp == NPW) {
    mma_warp<BN, SG>(ctl, tiles, tmem, u0, u1, nkb, o_tiles, a.S, x_shared);
  } else {
    int it = 0, seg = 0;
    for (int u = u0; u < u1; ++seg) {
      const int tile = u / nkb, kb0 = u - tile * nkb, kb1 = min(nkb, kb0 + (u1 - u));
      const int sgi = tile / o_tiles, ot = tile - sgi * o_tiles;
      const int s0 = sgi * SG, ns = min(SG, a.S - s0);
      const int64_t o0 = (int64_t)ot * BN;
      const bool first = kb0 == 0;   // this segment owns the bias (value and log-prob terms)
      if (tid < BN) {
        const int64_t o = o0 + tid;
#pragma unroll
        for (int s = 0; s < SG; ++s) {
          float bv = 0.0f;
          if (first && s < ns && o < a.out) {
            float sg, ep;
            bias_elem(a, s0 + s, o, sample, kLogProb, bv, sg, ep);
            if (kLogProb) { lp[s] += logp_elem(a.prior, bv); lq[s] += logq_elem(sg, ep); }
          }
          bias_s[s][tid] = bv;
        }
      }
      // Per-thread constants of the segment.  Thread t samples the weight quad (row t/8, chunk t%8) of every unit and
      // stages the activation chunks (rows t/8 + 32 j, chunk t%8): the shared-memory slots are the same in every
      // stage and the global addresses advance by 32 floats per unit, so the unit loop carries no index arithmetic.
      const int64_t o = o0 + wrow;
      const bool o_ok = tid < BN * 8 && o < a.out;
      const int64_t e_row = o * a.in + chunk * 4;
      float4 nmu = make_float4(0.f, 0.f, 0.f, 0.f), nrho = nmu;
      if (o_ok && kb0 * BK + chunk * 4 < a.in) {
        nmu = __ldg(reinterpret_cast<const float4 *>(a.w_mu + e_row + (int64_t)kb0 * BK));
        if (sample || kLogProb) nrho = __ldg(reinterpret_cast<const float4 *>(a.w_rho + e_row + (int64_t)kb0 * BK));
      }
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int stage = it % NS;
        if (it >= NS) mbar_wait(smem_u32(&ctl.empty[stage]), (uint32_t)(((it / NS) - 1) & 1));
        uint8_t *As = tiles + stage * R::kStage, *Bs = As + SG * A_TILE;
        const int kc = kb * BK + chunk * 4;
        const bool col_ok = kc < a.in;
        // activation loads are issued first and consumed after the sampling below has covered their latency
        float4 xv[SG][4];
#pragma unroll
        for (int s = 0; s < SG; ++s) {
          if (s < ns && (s == 0 || !x_shared)) {
            const float *xs = a.x + (int64_t)(s0 + s) * a.x_sstride + kc;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const bool ok = col_ok && xrow_ok[j];
              const float4 v = __ldg(reinterpret_cast<const float4 *>(ok ? xs + xrow_off[j] : a.x));
              xv[s][j] = ok ? v : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        }
        // this unit's mu/rho arrived during the previous unit; the next unit's are requested now
        const float4 cmu = nmu, crho = nrho;
        if (o_ok && kb + 1 < kb1 && kc + BK < a.in) {
          nmu = __ldg(reinterpret_cast<const float4 *>(a.w_mu + e_row + (int64_t)(kb + 1) * BK));
          if (sample || kLogProb) nrho = __ldg(reinterpret_cast<const float4 *>(a.w_rho + e_row + (int64_t)(kb + 1) * BK));
        }
        // weights: [BN out rows][32 k] per sample, formed in registers
        if (tid < BN * 8) {
          if (o_ok && col_ok) {
            const int64_t e = e_row + (int64_t)kb * BK;
            Quad q;
            q.mu[0] = cmu.x; q.mu[1] = cmu.y; q.mu[2] = cmu.z; q.mu[3] = cmu.w;
            q.rho[0] = crho.x; q.rho[1] = crho.y; q.rho[2] = crho.z; q.rho[3] = crho.w;
            if (sample || kLogProb) {
#pragma unroll
              for (int c = 0; c < 4; ++c) q.sg[c] = softplus_fast(q.rho[c]);
            }
            float lsg = 0.0f;
            if (kLogProb) lsg = logsigma_quad_fast(q.sg);
#pragma unroll
            for (int s = 0; s < SG; ++s) {
              if (s < ns) {
                float ep[4], w[4];
                sample_quad(a, s0 + s, e, q, sample, ep, w);
                *reinterpret_cast<float4 *>(Bs + s * R::kB + w_off) =
                    make_float4(to_tf32(w[0]), to_tf32(w[1]), to_tf32(w[2]), to_tf32(w[3]));
                if (kLogProb) {
                  lp[s] += logp_quad_fast(a.prior, w);
                  lq[s] += -4.0f * kHalfLog2Pi - lsg -
                           0.5f * (ep[0] * ep[0] + ep[1] * ep[1] + ep[2] * ep[2] + ep[3] * ep[3]);
                }
              }
            }
          } else {
#pragma unroll
            for (int s = 0; s < SG; ++s)
              if (s < ns) *reinterpret_cast<float4 *>(Bs + s * R::kB + w_off) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        // activations go in as they are: the tensor core reads the upper 19 bits of each fp32 (TF32 by truncation)
#pragma unroll
        for (int s = 0; s < SG; ++s) {
          if (s < ns && (s == 0 || !x_shared)) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<float4 *>(As + s * A_TILE + w_off + 4096 * j) = relu ? relu4(xv[s][j]) : xv[s][j];
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&ctl.full[stage]));
      }
      // this segment's accumulators are final: TMEM -> swizzled smem (the ring, idle now) -> y
      mbar_wait_parked(smem_u32(&ctl.acc), (uint32_t)(seg & 1));
      tc_fence_after_sync();
      bar_producers();  // bias_s
      drain_tile<BN, SG, 1>(tiles, tmem, ns, true, a.y + (int64_t)s0 * a.B * a.out, a.B * a.out, 0, a.B, o0, a.out,
                            a.vec_out, !(kb0 == 0 && kb1 == nkb), 1.0f, [&](int s, int c) { return bias_s[s][c]; });
      tc_fence_before_sync();
      bar_producers();  // ring and bias_s are reused by the next segment
      u += kb1 - kb0;
    }
  }
  __syncthreads();
  if (kLogProb) {  // the launcher keeps a log-prob launch inside one sample group, so lp[s] belongs to sample s
    const int ns_all = min(SG, a.S);
#pragma unroll
    for (int s = 0; s < SG; ++s)
      if (s < ns_all) block_sum2_atomic(lp[s], lq[s], red, a.logp + s, a.logq + s);
  }
  ctl2_teardown(ctl, R::kTmemCols);
}

// ==================================================================================================
// dgrad: dx_s[b][i] += sum_{o in segment} dz_s[b][o] W_s[o][i];  A = dz (K = o), B[i][o] = W^T
// ==================================================================================================
template <int BN, int SG>
__global__ void __launch_bounds__(NT2, kCtaPerSm) dgrad_sk_kernel(const LinArgs a_in, int nkb, int i_tiles, int total) {
  using R = Ring<BN, SG>;
  extern __shared__ uint8_t dsm[];
  __shared__ Ctl2 ctl;
  LinArgs a = a_in;
  rng_resolve(a.rng);
  uint8_t *tiles = align1024(dsm);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool sample = a.flags & BBB_F_SAMPLE;
  int u0, u1;
  unit_range(total, u0, u1);

  ctl2_setup(ctl, R::kTmemCols);
  const uint32_t tmem = ctl.tmem_base;
  const float osc = ((a.flags & BBB_F_SCALE_DX) && a.out_scale_dev) ? __ldg(a.out_scale_dev) : 1.0f;

  if (warp == NPW) {
    mma_warp<BN, SG>(ctl, tiles, tmem, u0, u1, nkb, i_tiles, a.S, false);
  } else {
    int it = 0, seg = 0;
    for (int u = u0; u < u1; ++seg) {
      const int tile = u / nkb, kb0 = u - tile * nkb, kb1 = min(nkb, kb0 + (u1 - u));
      const int sgi = tile / i_tiles, itl = tile - sgi * i_tiles;
      const int s0 = sgi * SG, ns = min(SG, a.S - s0);
      const int64_t i0 = (int64_t)itl * BN;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int stage = it % NS;
        if (it >= NS) mbar_wait(smem_u32(&ctl.empty[stage]), (uint32_t)(((it / NS) - 1) & 1));
        uint8_t *As = tiles + stage * R::kStage, *Bs = As + SG * A_TILE;
        const int64_t obase = (int64_t)kb * BK;
        // dz tile [128 b][32 o] per sample (mask = ReLU of this layer's own output): loads first
        float4 zv[SG][4];
#pragma unroll
        for (int s = 0; s < SG; ++s) {
          if (s < ns) {
            const int64_t base = (int64_t)(s0 + s) * a.B * a.out;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int idx = tid + NTP * j, row = idx >> 3;
              const int64_t o = obase + (idx & 7) * 4;
              float4 v = ld_row4(a.dy + base, row, o, a.B, a.out, a.vec_out);
              if (a.mask) v = mask4(v, ld_row4(a.mask + base, row, o, a.B, a.out, a.vec_out));
              zv[s][j] = v;
            }
          }
        }
        // W^T tile [BN i rows][32 o]: each thread samples one weight quad (o, 4 i) and scatters it transposed
        for (int idx = tid; idx < (BN / 4) * BK; idx += NTP) {
          const int iq = idx % (BN / 4), ol = idx / (BN / 4);
          const int64_t o = obase + ol, i = i0 + iq * 4;
          float w[SG][4];
          if (o < a.out && i < a.in) {
            const int64_t e = o * a.in + i;
            Quad q;
            load_quad(a, e, sample, q);
#pragma unroll
            for (int s = 0; s < SG; ++s) {
              float ep[4];
              if (s < ns) sample_quad(a, s0 + s, e, q, sample, ep, w[s]);
            }
          } else {
#pragma unroll
            for (int s = 0; s < SG; ++s)
#pragma unroll
              for (int c = 0; c < 4; ++c) w[s][c] = 0.0f;
          }
#pragma unroll
          for (int s = 0; s < SG; ++s) {
            if (s < ns) {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                *reinterpret_cast<float *>(Bs + s * R::kB + sw128_off(iq * 4 + c, ol >> 2) + (ol & 3) * 4) =
                    to_tf32(w[s][c]);
            }
          }
        }
#pragma unroll
        for (int s = 0; s < SG; ++s) {
          if (s < ns) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int idx = tid + NTP * j;
              const float4 v = zv[s][j];
              st_tile4(As + s * A_TILE, idx >> 3, idx & 7, v.x, v.y, v.z, v.w);
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&ctl.full[stage]));
      }
      mbar_wait_parked(smem_u32(&ctl.acc), (uint32_t)(seg & 1));
      tc_fence_after_sync();
      drain_tile<BN, SG, 1>(tiles, tmem, ns, true, a.dx + (int64_t)s0 * a.B * a.in, a.B * a.in, 0, a.B, i0, a.in,
                            a.vec_in, !(kb0 == 0 && kb1 == nkb), osc, [](int, int) { return 0.0f; },
                            (a.flags & BBB_F_DX_PREACT) ? a.x + (int64_t)s0 * a.x_sstride : nullptr);
      tc_fence_before_sync();
      bar_producers();
      u += kb1 - kb0;
    }
  }
  ctl2_teardown(ctl, R::kTmemCols);
}

// ==================================================================================================
// wgrad: G_s[i][o] = sum_b x_s[b][i] dz_s[b][o] for a 16-aligned window of BN columns, then for this CTA's own
// columns [o_lo, o_hi) and rows [i0, i0 + T):
//   t = G - gp w R(w);  grad_mu += t;  grad_rho += sigmoid(rho) (t eps - gq / sigma)      (eps regenerated)
// ==================================================================================================
struct RawQuad {
  float4 mu, rho;
};
__device__ __forceinline__ RawQuad load_raw(const LinArgs &a, int64_t e) {
  RawQuad r;
  r.mu = __ldg(reinterpret_cast<const float4 *>(a.w_mu + e));
  r.rho = __ldg(reinterpret_cast<const float4 *>(a.w_rho + e));
  return r;
}

template <int BN, int SG>
__global__ void __launch_bounds__(NT, kCtaPerSm) wgrad_bal_kernel(const LinArgs a_in, int T) {
  using SM = Smem<BN, SG>;
  static_assert(SG * BN * BM * 4 <= SM::kTiles, "G staging must fit in the operand buffers");
  extern __shared__ uint8_t dsm[];
  __shared__ Ctl ctl;
  LinArgs a = a_in;
  rng_resolve(a.rng);
  uint8_t *tiles = align1024(dsm);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t i0 = (int64_t)blockIdx.y * T, i_end = min(a.in, i0 + (int64_t)T);
  const int64_t o_lo = (int64_t)blockIdx.x * a.out / gridDim.x, o_hi = (int64_t)(blockIdx.x + 1) * a.out / gridDim.x;
  const int64_t o0 = (o_lo / 16) * 16;              // MMA window [o0, o0 + BN) covers [o_lo, o_hi)
  const bool sample = a.flags & BBB_F_SAMPLE, relu = a.flags & BBB_F_RELU_IN;
  const bool bias_cta = blockIdx.y == 0;
  const int nkb = (int)((a.B + BK - 1) / BK);
  constexpr uint32_t idesc = idesc_tf32(BM, BN);
  const float osc = a.out_scale_dev ? __ldg(a.out_scale_dev) : 1.0f;
  const bool accum_flag = a.flags & BBB_F_ACCUM;

  ctl_setup(ctl, SM::kTmemCols);
  const uint32_t tmem = ctl.tmem_base;
  float *Gs = reinterpret_cast<float *>(tiles);  // [SG][BN][128] after the MMAs of a group have completed
  __shared__ float colsum_s[SG][BN];             // sum_b dz[b][o] (bias gradients), first row tile only

  const int nrow = (int)(o_hi - o_lo), tq = (int)((i_end - i0) >> 2), nquad = nrow * tq;
  const int ol0 = (int)(o_lo - o0);

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
          uint8_t *Tt = As + s * A_TILE;
          st_tile4(Tt, iq * 4 + 0, bq, v[0].x, v[1].x, v[2].x, v[3].x);
          st_tile4(Tt, iq * 4 + 1, bq, v[0].y, v[1].y, v[2].y, v[3].y);
          st_tile4(Tt, iq * 4 + 2, bq, v[0].z, v[1].z, v[2].z, v[3].z);
          st_tile4(Tt, iq * 4 + 3, bq, v[0].w, v[1].w, v[2].w, v[3].w);
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
            uint8_t *Tt = Bs + s * SM::kB;
            st_tile4(Tt, oq * 4 + 0, bq, v[0].x, v[1].x, v[2].x, v[3].x);
            st_tile4(Tt, oq * 4 + 1, bq, v[0].y, v[1].y, v[2].y, v[3].y);
            st_tile4(Tt, oq * 4 + 2, bq, v[0].z, v[1].z, v[2].z, v[3].z);
            st_tile4(Tt, oq * 4 + 3, bq, v[0].w, v[1].w, v[2].w, v[3].w);
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
    // the first weight quads of the epilogue are fetched while the tensor pipe finishes
    int idx = tid;
    RawQuad nxt{};
    if (idx < nquad) {
      const int r = idx / tq;
      nxt = load_raw(a, (o_lo + r) * a.in + i0 + (int64_t)(idx - r * tq) * 4);
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
    // analytic epilogue over this CTA's (o, i quad) range; the next quad's mu/rho are in flight while one is computed
    while (idx < nquad) {
      const RawQuad cur = nxt;
      const int r = idx / tq, iq = idx - r * tq, ol = ol0 + r;
      const int64_t e = (o_lo + r) * a.in + i0 + (int64_t)iq * 4;
      idx += NT;
      if (idx < nquad) {
        const int rn = idx / tq;
        nxt = load_raw(a, (o_lo + rn) * a.in + i0 + (int64_t)(idx - rn * tq) * 4);
      }
      Quad q;
      q.mu[0] = cur.mu.x; q.mu[1] = cur.mu.y; q.mu[2] = cur.mu.z; q.mu[3] = cur.mu.w;
      q.rho[0] = cur.rho.x; q.rho[1] = cur.rho.y; q.rho[2] = cur.rho.z; q.rho[3] = cur.rho.w;
      float gm[4] = {0.f, 0.f, 0.f, 0.f}, gr[4] = {0.f, 0.f, 0.f, 0.f}, sgm[4], isg[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        q.sg[j] = softplus_f(q.rho[j]);
        sgm[j] = sigmoid_fast(q.rho[j]);
        isg[j] = __fdividef(1.0f, q.sg[j]);
      }
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
    // bias gradients: column sums of dz, by the CTAs of the first row tile, for their own columns
    if (bias_cta && tid < nrow) {
      const int64_t o = o_lo + tid;
      float gbm = 0.0f, gbr = 0.0f;
      for (int s = 0; s < ns; ++s) {
        const float colsum = colsum_s[s][ol0 + tid];
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
    __syncthreads();  // Gs (operand buffers) is reused by the next group
  }
  ctl_teardown(ctl, SM::kTmemCols);
}

// ==================================================================================================
// host-side launch logic
// ==================================================================================================
inline int cdiv_i(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

template <class K>
int set_smem(K kernel, int bytes) {
  BBB_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return BBB_OK;
}

inline int grid_for(int total_units) { return total_units < kCtaPerSm * kSMs ? total_units : kCtaPerSm * kSMs; }

template <int BN, int SG>
int launch_fwd(const LinArgs &a, cudaStream_t st) {
  const int o_tiles = cdiv_i(a.out, BN), nkb = cdiv_i(a.in, BK), s_groups = cdiv_i(a.S, SG);
  const bool lpq = a.flags & BBB_F_LOGPROB;
  const int smem = Ring<BN, SG>::kDyn;
  if (!(a.flags & BBB_F_OUT_ZEROED)) {
    BBB_CHECK_CUDA(cudaMemsetAsync(a.y, 0, sizeof(float) * (size_t)a.S * a.B * a.out, st));
    note_launch();
  }
  // one launch per sample group when log-probs are accumulated (a CTA's partial sums then belong to one group)
  const int launches = (lpq && s_groups > 1) ? s_groups : 1;
  for (int l = 0; l < launches; ++l) {
    LinArgs b = a;
    int groups = s_groups;
    if (launches > 1) {
      b.S = (l + 1 < launches) ? SG : a.S - l * SG;
      b.rng.sample_base += (uint32_t)(l * SG);
      if (b.x_sstride) b.x += (int64_t)l * SG * a.x_sstride;
      if (b.eps_w) { b.eps_w += (int64_t)l * SG * a.out * a.in; b.eps_b += (int64_t)l * SG * a.out; }
      b.y += (int64_t)l * SG * a.B * a.out;
      b.logp += l * SG; b.logq += l * SG;
      groups = 1;
    }
    const int total = groups * o_tiles * nkb;
    if (lpq) {
      if (int r = set_smem(fwd_sk_kernel<BN, SG, true>, smem)) return r;
      fwd_sk_kernel<BN, SG, true><<<grid_for(total), NT2, smem, st>>>(b, nkb, o_tiles, total);
    } else {
      if (int r = set_smem(fwd_sk_kernel<BN, SG, false>, smem)) return r;
      fwd_sk_kernel<BN, SG, false><<<grid_for(total), NT2, smem, st>>>(b, nkb, o_tiles, total);
    }
    BBB_CHECK_LAUNCH();
  }
  return BBB_OK;
}

template <int BN, int SG>
int launch_dgrad(const LinArgs &a, cudaStream_t st) {
  const int i_tiles = cdiv_i(a.in, BN), nkb = cdiv_i(a.out, BK), s_groups = cdiv_i(a.S, SG);
  const int total = s_groups * i_tiles * nkb;
  const int smem = Ring<BN, SG>::kDyn;
  if (!(a.flags & BBB_F_OUT_ZEROED)) {
    BBB_CHECK_CUDA(cudaMemsetAsync(a.dx, 0, sizeof(float) * (size_t)a.S * a.B * a.in, st));
    note_launch();
  }
  if (int r = set_smem(dgrad_sk_kernel<BN, SG>, smem)) return r;
  dgrad_sk_kernel<BN, SG><<<grid_for(total), NT2, smem, st>>>(a, nkb, i_tiles, total);
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

template <int BN, int SG>
int launch_wgrad(const LinArgs &a, cudaStream_t st) {
  const int n_it = cdiv_i(a.in, BM);
  const int T = ((cdiv_i(a.in, n_it) + 3) / 4) * 4;           // equal row tiles, a multiple of the Philox quad
  int n_c = (kCtaPerSm * kSMs) / n_it;
  if (n_c > a.out) n_c = (int)a.out;
  if (n_c < 1) n_c = 1;
  if (cdiv_i(a.out, n_c) + 15 > BN) n_c = cdiv_i(a.out, BN - 15);  // every column range must fit its MMA window
  dim3 grid(n_c, cdiv_i(a.in, T));
  const int smem = Smem<BN, SG>::kDyn;
  if (int r = set_smem(wgrad_bal_kernel<BN, SG>, smem)) return r;
  wgrad_bal_kernel<BN, SG><<<grid, NT, smem, st>>>(a, T);
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

}  // namespace

bool linear_sk_supported(const LinArgs &a) {
  return a.vec_in && a.in >= 4 && a.out >= 1 && a.B >= 1 && a.B <= BM && a.S >= 1;
}

int launch_linear_fwd_sk(const LinArgs &a, cudaStream_t st) {
  const bool two = a.S >= 2;
  if (a.out <= 16) return two ? launch_fwd<16, 2>(a, st) : launch_fwd<16, 1>(a, st);
  return two ? launch_fwd<32, 2>(a, st) : launch_fwd<32, 1>(a, st);
}

int launch_linear_bwd_sk(const LinArgs &a, cudaStream_t st) {
  const bool two = a.S >= 2;
  if (!(a.flags & BBB_F_NO_DX)) {
    int r;
    if (a.in <= 16) r = two ? launch_dgrad<16, 2>(a, st) : launch_dgrad<16, 1>(a, st);
    else r = two ? launch_dgrad<32, 2>(a, st) : launch_dgrad<32, 1>(a, st);
    if (r) return r;
  }
  if (a.flags & BBB_F_NO_WGRAD) return BBB_OK;
  if (a.out <= 16) return two ? launch_wgrad<16, 2>(a, st) : launch_wgrad<16, 1>(a, st);
  return two ? launch_wgrad<64, 2>(a, st) : launch_wgrad<64, 1>(a, st);
}

}  // namespace bbb
