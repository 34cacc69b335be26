// Weight-sampling Bayesian linear layer, tcgen05 kind::tf32, for batches of at most 128 rows (one UMMA M tile):
// the regime of BASELINE.json configs 1-4, where the contraction is tiny and the kernel is bound by the
// sampling work per weight (Philox + Box-Muller + prior/posterior terms), not by the tensor pipe.
//
// What decides the speed here is how evenly that per-weight work is spread over the 148 SMs and whether every
// SM has enough warps in flight to issue it, so the three kernels are organised around that:
//
//   fwd          stream-K.  The weight matrix is cut into units of [BN rows x 32 k]; the units, in (tile, k)
//                order, are dealt out in equal contiguous ranges to 2 x 148 CTAs (two co-resident CTAs per SM,
//                18 warps).  A CTA accumulates each run of units that belongs to one output tile in TMEM and
//                adds that partial tile to the (zero-filled) output with red.global.add.v4.f32.
//                Warp-specialised: 8 producer warps sample W_s = mu + sigma eps into the SWIZZLE_128B operand
//                ring (full/empty mbarriers, no CTA-wide barrier in the loop), 1 warp issues tcgen05.mma.
// The backward of this regime is one fused kernel per layer: bbb_linear_bwd_fused.cu.
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
  __shared__ int last_s0;
  LinArgs a = a_in;
  uint8_t *tiles = align1024(dsm);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) last_s0 = -1;      // (ordered before its use by the barrier in ctl2_setup)
  const bool sample = a.flags & BBB_F_SAMPLE, relu = a.flags & BBB_F_RELU_IN, x_shared = a.x_sstride == 0;
  int u0, u1;
  unit_range(total, u0, u1);

  ctl2_setup(ctl, R::kTmemCols);
  const uint32_t tmem = ctl.tmem_base;
  pdl_wait();              // everything above is local; from here on global memory of earlier kernels is read
  rng_resolve(a.rng);

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
    // A CTA's unit range may run through more than one sample group (one launch covers all groups): the registers
    // hold the log-prob sums of group `lp_s0`; they are handed over -- warp sums, one fp64 atomic per CTA and value --
    // when the group changes and after the last segment.
    int lp_s0 = -1;
    auto flush_logprob = [&]() {
#pragma unroll
      for (int s = 0; s < SG; ++s) {
        const float p = warp_sum(lp[s]), q = warp_sum(lq[s]);
        if (lane == 0) { red[s * 16 + warp] = p; red[s * 16 + 8 + warp] = q; }
        lp[s] = lq[s] = 0.0f;
      }
      bar_producers();
      if (tid < 2 * SG) {
        const int s = tid >> 1, which = tid & 1;
        double v = 0.0;
#pragma unroll
        for (int w8 = 0; w8 < NPW; ++w8) v += (double)red[s * 16 + which * 8 + w8];
        if (lp_s0 + s < a.S) atomicAdd((which ? a.logq : a.logp) + lp_s0 + s, v);
      }
      bar_producers();
    };
    int it = 0, seg = 0;
    for (int u = u0; u < u1; ++seg) {
      const int tile = u / nkb, kb0 = u - tile * nkb, kb1 = min(nkb, kb0 + (u1 - u));
      const int sgi = tile / o_tiles, ot = tile - sgi * o_tiles;
      const int s0 = sgi * SG, ns = min(SG, a.S - s0);
      if (kLogProb && s0 != lp_s0) {
        if (lp_s0 >= 0) flush_logprob();
        lp_s0 = s0;
      }
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
    if (kLogProb && tid == 0) last_s0 = lp_s0;
  }
  pdl_launch_dependents();   // the main loop is done: let the next kernel of the chain become resident
  __syncthreads();
  if (kLogProb && last_s0 >= 0) {  // the last group's sums (the MMA warp holds zeros), after the dependents were let in
    const int ns_last = min(SG, a.S - last_s0);
#pragma unroll
    for (int s = 0; s < SG; ++s)
      if (s < ns_last) block_sum2_atomic(lp[s], lq[s], red, a.logp + last_s0 + s, a.logq + last_s0 + s);
  }
  ctl2_teardown(ctl, R::kTmemCols);
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

inline int grid_for(int total_units) { return total_units < kCtaPerSm * sm_count() ? total_units : kCtaPerSm * sm_count(); }

template <int BN, int SG>
int launch_fwd(const LinArgs &a, cudaStream_t st) {
  const int o_tiles = cdiv_i(a.out, BN), nkb = cdiv_i(a.in, BK), s_groups = cdiv_i(a.S, SG);
  const bool lpq = a.flags & BBB_F_LOGPROB;
  const int smem = Ring<BN, SG>::kDyn;
  if (!(a.flags & BBB_F_OUT_ZEROED)) {
    BBB_CHECK_CUDA(cudaMemsetAsync(a.y, 0, sizeof(float) * (size_t)a.S * a.B * a.out, st));
    note_launch();
  }
  const int launches = 1;
  for (int l = 0; l < launches; ++l) {
    LinArgs b = a;
    const int groups = s_groups;
    const int total = groups * o_tiles * nkb;
    if (lpq) {
      if (int r = set_smem(fwd_sk_kernel<BN, SG, true>, smem)) return r;
      BBB_CHECK_CUDA(launch_pdl(fwd_sk_kernel<BN, SG, true>, dim3(grid_for(total)), dim3(NT2), smem, st, b, nkb, o_tiles, total));
    } else {
      if (int r = set_smem(fwd_sk_kernel<BN, SG, false>, smem)) return r;
      BBB_CHECK_CUDA(launch_pdl(fwd_sk_kernel<BN, SG, false>, dim3(grid_for(total)), dim3(NT2), smem, st, b, nkb, o_tiles, total));
    }
    BBB_CHECK_LAUNCH();
  }
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

}  // namespace bbb
