// Weight-sampling Bayesian linear layer, forward, for batches of at most 128 rows -- round-2 organisation.
//
//   D^T[o][b] = sum_k W_s[o][k] x_s[b][k]        (tcgen05 kind::tf32, M = 128 weight rows, N = 128 batch rows)
//
// The operands are swapped with respect to csrc/bbb_linear_sk.cu: the sampled weights are the A operand, so ONE
// activation tile [128 batch rows x 32 k] serves 4096 weights (it served 1024), and it is not staged by threads at all:
//   * warp 0 (one lane) is the TMA producer: per 32-wide k block it loads the mu and rho tiles [128 o x 32 k] (plain
//     row layout, read back by the samplers with conflict-free 16-byte shared loads) and the activation tile(s)
//     [128 b x 32 k] (SWIZZLE_128B = the K-major UMMA operand layout) with cp.async.bulk.tensor; out-of-range rows and
//     columns arrive as zeros, so the kernel carries no load predicates and no address arithmetic;
//   * warps 2..17 (512 threads) are the samplers: sigma once per weight, then per Monte-Carlo sample Philox4x32-10 ->
//     Box-Muller -> w = mu + sigma eps, the log-prior / log-posterior terms, and one 16-byte store of the TF32-rounded
//     quad into the SWIZZLE_128B weight tile.  Two quads x SG samples per thread and stage: straight-line code;
//   * warp 1 (one lane) issues the MMAs and releases stages with tcgen05.commit.
// Work split: the weight matrix is cut into (sample group, tile of T_o <= 128 output rows) pairs and every pair's k blocks
// are divided over its share of the one-CTA-per-SM grid, so a CTA owns ONE accumulator segment.  It drains the segment
// TMEM -> registers -> a plain [128 b][128 o] shared-memory tile (the operand ring is idle by then) and hands it to the
// TMA engine, which ADDS it to the zero-filled pre-activation output in L2 (cp.reduce.async.bulk.tensor): the split-K
// combine costs no per-thread atomics and no address arithmetic.  The pair's first CTA adds the sampled bias to its
// partial tile, so the output is the complete pre-activation once every CTA has finished -- no completion protocol.
// ReLU belongs to the consumer: when the input is a hidden layer's pre-activation (BBB_F_RELU_IN) the samplers apply
// max(x, 0) IN PLACE to the activation tiles the TMA has just loaded (two 16-byte chunks per thread and tile, ~1 % of the
// sampling work); the tensor core reads them as TF32 by truncation.
#include "bbb_tc_tiles.cuh"
#include "bbb_tma.cuh"
#include "bbb_mlp.h"

namespace bbb {
namespace {

using namespace tc;

constexpr int kSamplerWarps = 16;
constexpr int kSamplers = kSamplerWarps * 32;     // 512
constexpr int kThreads = 64 + kSamplers;          // + TMA warp + MMA warp
constexpr int TILE = 128 * 128;                   // bytes of one [128 rows][32 fp32] tile

template <int SG>
struct FwdCfg {
  static constexpr int kStages = SG == 2 ? 2 : 3;
  static constexpr int kStage = (2 + 2 * SG) * TILE;       // mu | rho | x[SG] | W[SG]
  static constexpr int kDyn = kStages * kStage + 1024;
  static constexpr uint32_t kTmemCols = SG * 128;
};

struct FwdCtl {
  uint64_t full_in[3], full_w[3], empty[3], acc_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ void bar_samplers() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_cta(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// (no memory clobber: the __threadfence() that follows the drain orders it, and volatile asm statements keep their order)
__device__ __forceinline__ void red_add_f32(const float *p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v));
}
// 16-byte shared-memory accesses by 32-bit shared-space address (no generic-address resolution in the inner loop)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// 16 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// MODE 0: eps from Philox, 1: eps injected from memory (parity mode), 2: w = mu (no sampling)
template <int SG, int MODE, bool kLogProb>
__global__ void __launch_bounds__(kThreads, 1)
ws_fwd_kernel(const __grid_constant__ CUtensorMap tm_mu, const __grid_constant__ CUtensorMap tm_rho,
              const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y, const MlpFwdArgs a) {
  using Cfg = FwdCfg<SG>;
  constexpr int NS = Cfg::kStages;
  extern __shared__ uint8_t dsm[];
  __shared__ FwdCtl ctl;
  __shared__ __align__(16) float bias_s[SG][128];
  __shared__ float red[2 * SG * kSamplerWarps];
  uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(dsm) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  unsigned long long *tl = tid == 64 ? a.timeline : nullptr;     // the first sampler thread stamps the phases
  stamp(tl, 0);

  // ---- this CTA's segment: (sample group, output-row tile) pair + a range of k blocks ---------------------------------
  int pair, part, cnt;
  {
    const int bid = blockIdx.x, big = a.rem * (a.base + 1);
    if (bid < big) { pair = bid / (a.base + 1); part = bid - pair * (a.base + 1); cnt = a.base + 1; }
    else { const int b2 = bid - big; pair = a.rem + b2 / a.base; part = b2 - (b2 / a.base) * a.base; cnt = a.base; }
  }
  const int g = pair / a.n_ot, ot = pair - g * a.n_ot;
  const int s0 = g * SG, ns = min(SG, a.S - s0);
  const int o0 = ot * a.T_o, rows = min(a.T_o, a.out - o0);
  const int kb0 = (int)((int64_t)part * a.nkb / cnt), kb1 = (int)((int64_t)(part + 1) * a.nkb / cnt);
  const bool need_rho = MODE != 2 || kLogProb;
  const int nx = a.x_shared ? 1 : ns;

  if (wid == 1) tmem_alloc(smem_u32(&ctl.tmem_base), Cfg::kTmemCols);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      mbar_init(smem_u32(&ctl.full_in[s]), 1);
      mbar_init(smem_u32(&ctl.full_w[s]), kSamplerWarps);
      mbar_init(smem_u32(&ctl.empty[s]), 1);
    }
    mbar_init(smem_u32(&ctl.acc_full), 1);
    mbar_fence_init();
    tma::prefetch_map(&tm_mu);
    if (need_rho) tma::prefetch_map(&tm_rho);
    tma::prefetch_map(&tm_x);
    tma::prefetch_map(&tm_y);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = ctl.tmem_base;
  stamp(tl, 1);
  // Programmatic dependent launch: everything up to here is local.  mu and rho are written by the optimiser only, which
  // completed before the previous kernel of the chain passed ITS wait, so the producer requests the first stages'
  // parameter tiles before waiting; activations (and everything else) are touched after pdl_wait().
  if (wid != 0) pdl_wait();
  stamp(tl, 2);

  float lp[SG], lq[SG];
#pragma unroll
  for (int s = 0; s < SG; ++s) lp[s] = lq[s] = 0.0f;

  if (wid == 0) {
    // ================================ TMA producer ====================================================================
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)((1 + (need_rho ? 1 : 0) + nx) * TILE);
      const int n_pre = min(NS, kb1 - kb0);
      for (int it = 0; it < n_pre; ++it) {              // parameter tiles of the first stages: before the wait
        const uint32_t sb = smem_u32(tiles + it * Cfg::kStage), bar = smem_u32(&ctl.full_in[it]);
        tma::arrive_expect_tx(bar, bytes);
        tma::load_2d(sb, &tm_mu, bar, (kb0 + it) * 32, o0);
        if (need_rho) tma::load_2d(sb + TILE, &tm_rho, bar, (kb0 + it) * 32, o0);
      }
      pdl_wait();
      int it = 0;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int stage = it % NS;
        const uint32_t sb = smem_u32(tiles + stage * Cfg::kStage), bar = smem_u32(&ctl.full_in[stage]);
        if (it >= NS) {
          mbar_wait(smem_u32(&ctl.empty[stage]), (uint32_t)(((it / NS) - 1) & 1));
          tma::arrive_expect_tx(bar, bytes);
          tma::load_2d(sb, &tm_mu, bar, kb * 32, o0);
          if (need_rho) tma::load_2d(sb + TILE, &tm_rho, bar, kb * 32, o0);
        }
#pragma unroll
        for (int s = 0; s < SG; ++s)
          if (s < nx) tma::load_3d(sb + (2 + s) * TILE, &tm_x, bar, kb * 32, 0, a.x_shared ? 0 : s0 + s);
      }
    } else {
      pdl_wait();
    }
    __syncwarp();
    pdl_launch_dependents();
  } else if (wid == 1) {
    // ================================ MMA issuer ======================================================================
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_tf32(128, 128);
      int it = 0;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int stage = it % NS;
        const uint32_t ph = (uint32_t)((it / NS) & 1);
        mbar_wait_parked(smem_u32(&ctl.full_in[stage]), ph);
        mbar_wait_parked(smem_u32(&ctl.full_w[stage]), ph);
        tc_fence_after_sync();
        const uint32_t sb = smem_u32(tiles + stage * Cfg::kStage);
#pragma unroll
        for (int s = 0; s < SG; ++s)
          if (s < ns)
            tcx::issue_block(tmem + s * 128, sb + (2 + SG + s) * TILE, sb + (2 + (a.x_shared ? 0 : s)) * TILE, idesc, kb == kb0);
        mma_commit(smem_u32(&ctl.empty[stage]));
      }
      mma_commit(smem_u32(&ctl.acc_full));
    }
    __syncwarp();
    pdl_launch_dependents();
  } else {
    // ================================ samplers ========================================================================
    RngDev rng = a.rng;
    rng_resolve(rng);
    const int st = tid - 64, c = st & 7, r0 = st >> 3;            // quads (row r0 + 64 j, 16-byte chunk c), j = 0, 1
    const uint32_t w_off = sw128_off(r0, c), p_off = (uint32_t)(r0 * 128 + c * 16);
    const bool row_ok[2] = {r0 < rows, r0 + 64 < rows};
    const uint32_t q_base[2] = {(uint32_t)(o0 + r0) * (uint32_t)a.in4 + (uint32_t)c,
                                (uint32_t)(o0 + r0 + 64) * (uint32_t)a.in4 + (uint32_t)c};
    // the tile's bias sample (needed by the finalisation at the very end: computed now, off the tail).  Its log-prob
    // terms are counted by the pair's first CTA only.
    if (st < 128 * SG) {
      const bool sample = MODE != 2;
      const int s = st >> 7, o_l = st & 127;
      float bv = 0.0f;
      if (s < ns && o_l < rows) {
        const int o = o0 + o_l;
        const float bmu = __ldg(a.b_mu + o);
        const float bsg = need_rho ? softplus_f(__ldg(a.b_rho + o)) : 0.0f;
        float ep = 0.0f;
        if (sample)
          ep = MODE == 1 ? __ldg(a.eps_b + (int64_t)(s0 + s) * a.out + o)
                         : philox_normal1(rng, rng.tensor_b, rng.sample_base + (uint32_t)(s0 + s), (uint64_t)o);
        bv = sample ? __fadd_rn(bmu, __fmul_rn(bsg, ep)) : bmu;
        if (kLogProb && part == 0) { lp[s] += logp_elem(a.prior, bv); lq[s] += logq_elem(bsg, ep); }
      }
      bias_s[s][o_l] = bv;
    }
    const uint32_t tiles_u32 = smem_u32(tiles);
    const bool relu_in = a.flags & BBB_F_RELU_IN;
    int it = 0;
    for (int kb = kb0; kb < kb1; ++kb, ++it) {
      const int stage = it % NS;
      mbar_wait(smem_u32(&ctl.full_in[stage]), (uint32_t)((it / NS) & 1));
      if (it < 4) stamp(tl, 3 + 2 * it);
      const uint32_t sb = tiles_u32 + (uint32_t)(stage * Cfg::kStage);
      const bool col_ok = kb * 32 + c * 4 < a.in;
      if (relu_in) {
        // the activation tiles hold the producer's PRE-activation: max(., 0) in place (elementwise, so the swizzle
        // does not matter: thread t owns bytes [16 t, 16 t + 16) of each 8 KB half of a tile)
#pragma unroll
        for (int s = 0; s < SG; ++s) {
          if (s >= nx) break;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t ad = sb + (2 + s) * TILE + (uint32_t)(st * 16 + h * 8192);
            const float4 v = lds128(ad);
            sts128(ad, fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (!row_ok[j]) {               // rows of the next tile: zero weights, so their accumulator lanes hold exact zeros
#pragma unroll
          for (int s = 0; s < SG; ++s)
            if (s < ns) sts128(sb + (2 + SG + s) * TILE + w_off + j * 8192, 0.f, 0.f, 0.f, 0.f);
          continue;
        }
        const float4 m4 = lds128(sb + p_off + j * 8192);
        const float mu[4] = {m4.x, m4.y, m4.z, m4.w};
        float sg[4] = {0.f, 0.f, 0.f, 0.f}, lsg = 0.0f;
        if (need_rho) {
          const float4 r4 = lds128(sb + TILE + p_off + j * 8192);
          sg[0] = softplus_fast(r4.x); sg[1] = softplus_fast(r4.y); sg[2] = softplus_fast(r4.z); sg[3] = softplus_fast(r4.w);
          if (kLogProb) lsg = logsigma_quad_fast(sg);
        }
#pragma unroll
        for (int s = 0; s < SG; ++s) {
          if (s >= ns) break;
          float ep[4] = {0.f, 0.f, 0.f, 0.f}, w[4];
          if (MODE == 0) {
            philox_normal4(rng, rng.tensor_w, rng.sample_base + (uint32_t)(s0 + s), q_base[j] + (uint32_t)kb * 8u, ep);
          } else if (MODE == 1) {
            if (col_ok) {
              const float4 e4 = __ldg(reinterpret_cast<const float4 *>(
                  a.eps_w + ((int64_t)(s0 + s) * a.out + o0 + r0 + 64 * j) * a.in + kb * 32 + c * 4));
              ep[0] = e4.x; ep[1] = e4.y; ep[2] = e4.z; ep[3] = e4.w;
            }
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) w[e] = MODE == 2 ? mu[e] : fmaf(sg[e], ep[e], mu[e]);
          sts128(sb + (2 + SG + s) * TILE + w_off + j * 8192, to_tf32(w[0]), to_tf32(w[1]), to_tf32(w[2]), to_tf32(w[3]));
          if (kLogProb && col_ok) {
            lp[s] += logp_quad_fast(a.prior, w);
            lq[s] += -4.0f * kHalfLog2Pi - lsg - 0.5f * (ep[0] * ep[0] + ep[1] * ep[1] + ep[2] * ep[2] + ep[3] * ep[3]);
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(smem_u32(&ctl.full_w[stage]));
      if (it < 4) stamp(tl, 4 + 2 * it);
    }
    pdl_launch_dependents();

    // ---- drain: TMEM [lane = o][column = b] -> shared staging [b][o] -> ONE bulk reduce-add per sample ----------------
    // The operand ring is idle now: stage s holds sample s's partial tile as a plain [128 b][128 o] fp32 matrix (a warp
    // writes 32 consecutive o of one batch row: conflict-free, immediate offsets), and the TMA engine adds it to the
    // zero-filled pre-activation scratch in L2 (cp.reduce.async.bulk.tensor .add): no per-thread atomics, no address
    // arithmetic; rows beyond the batch and columns beyond `out` are dropped by the TMA, accumulator lanes beyond the
    // tile's rows hold exact zeros (zero weight rows above).
    mbar_wait(smem_u32(&ctl.acc_full), 0u);
    tc_fence_after_sync();
    stamp(tl, 11);
    {
      const int q = wid & 3, cg = (wid - 2) >> 2;          // TMEM lane quarter of this warp, 32-column group
      const uint32_t srow = tiles_u32 + (uint32_t)((q * 32 + lane) * 4 + cg * 32 * 512);
#pragma unroll
      for (int s = 0; s < SG; ++s) {
        if (s >= ns) break;
        const float bv = part == 0 ? bias_s[s][q * 32 + lane] : 0.0f;     // the pair's first CTA carries the bias
#pragma unroll
        for (int cb = 0; cb < 32; cb += 16) {
          float v[16];
          tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 128 + cg * 32 + cb), v);
#pragma unroll
          for (int jj = 0; jj < 16; ++jj)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(srow + (uint32_t)(s * Cfg::kStage + (cb + jj) * 512)), "f"(v[jj] + bv) : "memory");
        }
      }
    }
    tc_fence_before_sync();
    fence_proxy_async_smem();
    bar_samplers();
    stamp(tl, 12);
    if (st == 0) {
#pragma unroll
      for (int s = 0; s < SG; ++s)
        if (s < ns) tma::reduce_add_3d(&tm_y, tiles_u32 + (uint32_t)(s * Cfg::kStage), o0, 0, s0 + s);
      tma::bulk_commit();
    }
    stamp(tl, 13);
    tc_fence_before_sync();
    // ---- log-prob sums of this CTA: warp sums -> one fp64 atomic per value (while the reduce-add is in flight) --------
    if (kLogProb) {
      const int sw = wid - 2;
#pragma unroll
      for (int s = 0; s < SG; ++s) {
        const float p = warp_sum(lp[s]), q2 = warp_sum(lq[s]);
        if (lane == 0) { red[(2 * s) * kSamplerWarps + sw] = p; red[(2 * s + 1) * kSamplerWarps + sw] = q2; }
      }
      bar_samplers();
      if (st >= 32 && st < 32 + 2 * SG) {
        const int s = (st - 32) >> 1, which = st & 1;
        if (s < ns) {
          double v = 0.0;
#pragma unroll
          for (int w8 = 0; w8 < kSamplerWarps; ++w8) v += (double)red[(2 * s + which) * kSamplerWarps + w8];
          atomicAdd((which ? a.logq : a.logp) + s0 + s, v);
        }
      }
    }
    // the partial tile has been added to the output (and its staging read).  Waiting only for the READ of the staging
    // tile (cp.async.bulk.wait_group.read) was measured: no gain -- the grid's completion waits for the adds anyway
    if (st == 0) tma::bulk_wait_all();
  }
  stamp(tl, 15);
  tc_fence_before_sync();
  __syncthreads();
  if (wid == 1) tmem_dealloc(tmem, Cfg::kTmemCols);
}

template <int SG, int MODE, bool kLogProb>
int launch_one(const CUtensorMap *tm, const MlpFwdArgs &a, int grid, cudaStream_t st) {
  auto kernel = ws_fwd_kernel<SG, MODE, kLogProb>;
  BBB_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdCfg<SG>::kDyn));
  BBB_CHECK_CUDA(launch_pdl(kernel, dim3(grid), dim3(kThreads), (size_t)FwdCfg<SG>::kDyn, st, tm[0], tm[1], tm[2], tm[3], a));
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

template <int SG>
int launch_sg(const CUtensorMap *tm, const MlpFwdArgs &a, int grid, int mode, bool lpq, cudaStream_t st) {
  if (mode == 0) return lpq ? launch_one<SG, 0, true>(tm, a, grid, st) : launch_one<SG, 0, false>(tm, a, grid, st);
  if (mode == 1) return lpq ? launch_one<SG, 1, true>(tm, a, grid, st) : launch_one<SG, 1, false>(tm, a, grid, st);
  return lpq ? launch_one<SG, 2, true>(tm, a, grid, st) : launch_one<SG, 2, false>(tm, a, grid, st);
}

inline int cdiv_i(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
inline bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

bool mlp_fwd_layer_supported(const MlpLayerDesc &l, int64_t S, int64_t B) {
  if (!(B >= 1 && B <= 128 && S >= 1 && l.in >= 4 && l.in % 4 == 0 && l.out >= 4 && l.out % 4 == 0)) return false;
  if (!(al16(l.x) && al16(l.w_mu) && al16(l.w_rho) && al16(l.eps_w) && al16(l.y))) return false;
  const int n_ot = cdiv_i(l.out, 128), groups = cdiv_i(S, 2);
  return (int64_t)n_ot * groups <= sm_count() && l.in * l.out / 4 < (int64_t)1 << 32;
}

// One layer: x [Sx,B,in] (x_shared: [B,in]; BBB_F_RELU_IN: max(x, 0) is applied to the loaded tiles) -> y [S,B,out] +=
// x W_s^T + b_s (y must be zero-filled), log-prob sums added to logp / logq.
int launch_mlp_fwd_layer(const MlpLayerDesc &l, int64_t S, int64_t B, const RngDev &rng, const PriorDev &prior, int flags,
                         double *logp, double *logq, cudaStream_t st) {
  const bool sample = flags & BBB_F_SAMPLE, lpq = flags & BBB_F_LOGPROB;
  const int mode = !sample ? 2 : (l.eps_w ? 1 : 0);
  MlpFwdArgs a{};
  a.b_mu = l.b_mu; a.b_rho = l.b_rho; a.eps_w = l.eps_w; a.eps_b = l.eps_b;
  a.logp = logp; a.logq = logq;
  a.rng = rng; a.prior = prior;
  a.S = (int)S; a.B = (int)B; a.in = (int)l.in; a.out = (int)l.out; a.in4 = (int)(l.in / 4);
  a.n_ot = cdiv_i(l.out, 128);
  a.T_o = ((cdiv_i(l.out, a.n_ot) + 3) / 4) * 4;
  a.n_ot = cdiv_i(l.out, a.T_o);
  a.nkb = cdiv_i(l.in, 32);
  a.flags = flags; a.x_shared = l.x_shared ? 1 : 0;
  a.timeline = debug_timeline();
  const int sg = S >= 2 ? 2 : 1, groups = cdiv_i(S, sg), pairs = groups * a.n_ot;
  int grid = sm_count();
  if ((int64_t)pairs * a.nkb < grid) grid = pairs * a.nkb;
  a.base = grid / pairs; a.rem = grid % pairs;
  CUtensorMap tm[4];
  if (int r = tma::make_map(&tm[0], l.w_mu, l.in, l.out, 0, 32, 128, tma::kNone)) return r;
  tm[1] = tm[0];
  if (l.w_rho)
    if (int r = tma::make_map(&tm[1], l.w_rho, l.in, l.out, 0, 32, 128, tma::kNone)) return r;
  if (int r = tma::make_map(&tm[2], l.x, l.in, B, l.x_shared ? 1 : S, 32, 128, tma::kSw128)) return r;
  if (int r = tma::make_map(&tm[3], l.y, l.out, B, S, 128, 128, tma::kNone)) return r;   // split-K reduce-add target
  return sg == 2 ? launch_sg<2>(tm, a, grid, mode, lpq, st) : launch_sg<1>(tm, a, grid, mode, lpq, st);
}

}  // namespace bbb
